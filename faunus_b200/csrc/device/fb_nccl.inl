// Replica exchange between GPUs without leaving C++ (parallel tempering): an NCCL communicator
// owned by the context, point-to-point ncclSend/ncclRecv on DEVICE buffers over NVLink / NVSwitch on the context's
// stream. Replaces the reference's MPI messages (src/move.cpp:860-923, src/mpicontroller.cpp:192-259):
//   fb_nccl_exchange_state   the packed state of the accepted slot (box, group sizes, x y z q id of every particle — the
//                            `ExchangeParticles` buffer, the group sizes and `exchangeVolume` in one message) goes
//                            mirror → partner's mirror; the host only receives a copy to keep its Space in step
//   fb_nccl_sendrecv_host    a few doubles (the 8-byte energy change), host ↔ host through pinned + device staging
//   fb_nccl_allgather_host   one double per rank (checkRandomEngineState; doubles as the barrier)
// libnccl is opened at run time (dlopen): the library itself has no link-time dependency on it, and a process that
// never tempers never loads it. Included at the end of fb_api.cu.
#include <dlfcn.h>
#include <nccl.h> // types and constants only

namespace {

struct NcclApi
{
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi& ncclApi()
{
    static NcclApi api;
    if (api.lib != nullptr) {
        return api;
    }
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
        api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib != nullptr) {
            break;
        }
    }
    if (api.lib == nullptr) {
        throw CudaError{std::string("cannot open libnccl: ") + dlerror()};
    }
    auto sym = [&](const char* name) {
        void* p = dlsym(api.lib, name);
        if (p == nullptr) {
            throw CudaError{std::string("libnccl lacks ") + name};
        }
        return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
    api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    return api;
}

void ncclCheck(ncclResult_t rc, const char* what)
{
    if (rc != ncclSuccess) {
        throw CudaError{std::string(what) + ": " + ncclApi().GetErrorString(rc)};
    }
}

ncclComm_t commOf(fb_ctx* c)
{
    if (c->nccl_comm == nullptr) {
        throw CudaError{"fb_nccl_init has not been called on this context"};
    }
    return static_cast<ncclComm_t>(c->nccl_comm);
}

/** device buffers `send` → partner, partner → `recv`, n doubles each, on the context's stream */
void ncclExchange(fb_ctx* c, const double* send, double* recv, size_t n, int partner)
{
    auto& api = ncclApi();
    if (partner < 0 || partner >= c->nccl_size || partner == c->nccl_rank) {
        throw CudaError{"bad exchange partner"};
    }
    ncclCheck(api.GroupStart(), "ncclGroupStart");
    ncclCheck(api.Send(send, n, ncclDouble, partner, commOf(c), c->stream), "ncclSend");
    ncclCheck(api.Recv(recv, n, ncclDouble, partner, commOf(c), c->stream), "ncclRecv");
    ncclCheck(api.GroupEnd(), "ncclGroupEnd");
}

} // namespace

FB_API int fb_nccl_unique_id(char out[128])
{
    static_assert(NCCL_UNIQUE_ID_BYTES == 128, "fb_nccl_unique_id hands out 128 bytes");
    try {
        if (out == nullptr) {
            throw CudaError{"null buffer"};
        }
        ncclUniqueId id;
        ncclCheck(ncclApi().GetUniqueId(&id), "ncclGetUniqueId");
        std::memcpy(out, id.internal, NCCL_UNIQUE_ID_BYTES);
        return FB_OK;
    }
    catch (const CudaError& e) {
        g_create_error = e.msg; // no context to hold the message: fb_last_error(NULL)
        return FB_ERR_INVALID;
    }
}

FB_API int fb_nccl_init(fb_ctx* c, const char id_bytes[128], int rank, int size)
{
    return guarded(c, [&] {
        if (c->nccl_comm != nullptr) {
            throw CudaError{"the context already has a communicator"};
        }
        if (!id_bytes || size < 2 || rank < 0 || rank >= size) {
            throw CudaError{"bad communicator arguments"};
        }
        CUDA_CHECK(cudaSetDevice(c->device));
        ncclUniqueId id;
        std::memcpy(id.internal, id_bytes, NCCL_UNIQUE_ID_BYTES);
        ncclComm_t comm = nullptr;
        ncclCheck(ncclApi().CommInitRank(&comm, size, id, rank), "ncclCommInitRank");
        c->nccl_comm = comm;
        c->nccl_rank = rank;
        c->nccl_size = size;
    });
}

FB_API int fb_nccl_finalize(fb_ctx* c)
{
    return guarded(c, [&] {
        if (c->nccl_comm != nullptr) {
            CUDA_CHECK(cudaStreamSynchronize(c->stream));
            ncclApi().CommDestroy(static_cast<ncclComm_t>(c->nccl_comm));
            c->nccl_comm = nullptr;
        }
    });
}

FB_API int fb_nccl_exchange_state(fb_ctx* c, int send_slot, int s, int partner, double* host_received)
{
    return guarded(c, [&] {
        flushPending(c);
        checkSlot(c, send_slot);
        checkSlot(c, s);
        const size_t n = fb_state_doubles(c);
        c->d_state.ensure(n);
        c->d_state_recv.ensure(n);
        c->h_state.ensure(n);
        packStateKernel<<<gridFor(c, c->n_slots, 256), 256, 0, c->stream>>>(makeView(c, send_slot), c->d_state.ptr);
        launched(c, "packStateKernel");
        ncclExchange(c, c->d_state.ptr, c->d_state_recv.ptr, n, partner);
        // the partner's state: into the mirror on the device, and a copy for the host's Space
        CUDA_CHECK(cudaMemcpyAsync(c->h_state.ptr, c->d_state_recv.ptr, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        importPackedState(c, s, c->d_state_recv.ptr, c->h_state.ptr); // synchronises the stream
        if (host_received != nullptr) {
            std::memcpy(host_received, c->h_state.ptr, n * sizeof(double));
        }
        c->bytes_exchanged += n * sizeof(double);
    });
}

FB_API int fb_nccl_sendrecv_host(fb_ctx* c, double* data, size_t n, int partner)
{
    return guarded(c, [&] {
        if (!data || n == 0) {
            throw CudaError{"empty message"};
        }
        c->d_xchg.ensure(2 * n);
        c->h_xchg.ensure(n);
        std::memcpy(c->h_xchg.ptr, data, n * sizeof(double));
        CUDA_CHECK(cudaMemcpyAsync(c->d_xchg.ptr, c->h_xchg.ptr, n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        ncclExchange(c, c->d_xchg.ptr, c->d_xchg.ptr + n, n, partner);
        CUDA_CHECK(cudaMemcpyAsync(c->h_xchg.ptr, c->d_xchg.ptr + n, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        std::memcpy(data, c->h_xchg.ptr, n * sizeof(double));
        c->bytes_exchanged += n * sizeof(double);
    });
}

FB_API int fb_nccl_allgather_host(fb_ctx* c, double value, double* out)
{
    return guarded(c, [&] {
        if (!out) {
            throw CudaError{"null buffer"};
        }
        const size_t size = static_cast<size_t>(c->nccl_size);
        c->d_xchg.ensure(1 + size);
        c->h_xchg.ensure(std::max<size_t>(size, 1));
        c->h_xchg.ptr[0] = value;
        CUDA_CHECK(cudaMemcpyAsync(c->d_xchg.ptr, c->h_xchg.ptr, sizeof(double), cudaMemcpyHostToDevice, c->stream));
        ncclCheck(ncclApi().AllGather(c->d_xchg.ptr, c->d_xchg.ptr + 1, 1, ncclDouble, commOf(c), c->stream), "ncclAllGather");
        CUDA_CHECK(cudaMemcpyAsync(c->h_xchg.ptr, c->d_xchg.ptr + 1, size * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        std::memcpy(out, c->h_xchg.ptr, size * sizeof(double));
    });
}

FB_API unsigned long long fb_nccl_bytes_exchanged(const fb_ctx* c)
{
    return c ? c->bytes_exchanged : 0ull;
}
