// Replica communicators for parallel tempering (the reference uses MPI point-to-point through MPL:
// src/move.cpp:863, 909; src/mpicontroller.cpp:216, 236, 257).
//
//  * CallbackComm — one process per replica / GPU: the exchange primitives are C callbacks supplied by
//    the launcher (faunus_b200/replica.py implements them over torch.distributed: NCCL send/recv
//    between GPUs over NVLink, or gloo on CPU for tests).
//  * LocalComm — all replicas in one process, one thread each (in-process rendezvous); used by tests and
//    by single-process multi-GPU runs where the 8-byte exchanges need no network at all.
#pragma once
#include "moves.hpp"
#include <condition_variable>
#include <mutex>

extern "C" {
/** Exchange primitives provided by the launcher; `user` is passed back verbatim */
typedef struct
{
    int rank;
    int size;
    void* user;
    void (*barrier)(void* user);
    void (*sendrecv_replace)(void* user, double* data, size_t n, int partner);
    void (*gather)(void* user, double value, double* out /*[size], valid on rank 0*/);
} fb_replica_callbacks;
}

namespace fb {

class CallbackComm : public ReplicaComm
{
    fb_replica_callbacks cb;

  public:
    explicit CallbackComm(const fb_replica_callbacks& callbacks)
        : cb(callbacks)
    {
        if (!cb.barrier || !cb.sendrecv_replace || !cb.gather || cb.size < 1 || cb.rank < 0 || cb.rank >= cb.size) {
            throw std::runtime_error("incomplete replica callbacks");
        }
    }
    int rank() const override { return cb.rank; }
    int size() const override { return cb.size; }
    void barrier() override { cb.barrier(cb.user); }
    void sendrecvReplace(double* data, size_t n, int partner) override { cb.sendrecv_replace(cb.user, data, n, partner); }
    std::vector<double> gather(double value) override
    {
        std::vector<double> out(static_cast<size_t>(cb.size), value);
        cb.gather(cb.user, value, out.data());
        return out;
    }
};

/** Shared rendezvous state of the replicas living in one process */
struct LocalExchange
{
    explicit LocalExchange(int size)
        : size(size)
        , mailbox(static_cast<size_t>(size))
        , posted(static_cast<size_t>(size), 0)
        , gathered(static_cast<size_t>(size), 0.0)
    {
    }
    int size;
    std::mutex mutex;
    std::condition_variable cv;
    std::vector<std::vector<double>> mailbox; //!< message addressed TO rank i
    std::vector<int> posted;
    std::vector<double> gathered;
    int barrier_count = 0;
    long barrier_generation = 0;
    bool failed = false;
};

class LocalComm : public ReplicaComm
{
    std::shared_ptr<LocalExchange> ex;
    int my_rank;

  public:
    /** optional: exchange the whole state in one go (the B200 build ships the packed device mirror) */
    std::function<bool(Space&, int, VolumeMethod, Change&)> state_exchanger;
    bool exchangeState(Space& spc, int partner, VolumeMethod method, Change& change) override
    {
        return state_exchanger ? state_exchanger(spc, partner, method, change) : false;
    }

  private:

    void wait(std::unique_lock<std::mutex>& lock, const std::function<bool()>& pred)
    {
        ex->cv.wait(lock, [&] { return ex->failed || pred(); });
        if (ex->failed) {
            throw std::runtime_error("another replica failed");
        }
    }

  public:
    LocalComm(std::shared_ptr<LocalExchange> exchange, int rank)
        : ex(std::move(exchange))
        , my_rank(rank)
    {
    }
    int rank() const override { return my_rank; }
    int size() const override { return ex->size; }
    void fail()
    {
        std::lock_guard<std::mutex> lock(ex->mutex);
        ex->failed = true;
        ex->cv.notify_all();
    }
    void barrier() override
    {
        std::unique_lock<std::mutex> lock(ex->mutex);
        const long generation = ex->barrier_generation;
        if (++ex->barrier_count == ex->size) {
            ex->barrier_count = 0;
            ex->barrier_generation++;
            ex->cv.notify_all();
        }
        else {
            wait(lock, [&] { return ex->barrier_generation != generation; });
        }
    }
    void sendrecvReplace(double* data, size_t n, int partner) override
    {
        std::unique_lock<std::mutex> lock(ex->mutex);
        wait(lock, [&] { return ex->posted[partner] == 0; }); // partner's mailbox free
        ex->mailbox[partner].assign(data, data + n);
        ex->posted[partner] = 1;
        ex->cv.notify_all();
        wait(lock, [&] { return ex->posted[my_rank] == 1; });
        if (ex->mailbox[my_rank].size() != n) {
            ex->failed = true;
            ex->cv.notify_all();
            throw std::runtime_error("replica message size mismatch");
        }
        std::copy(ex->mailbox[my_rank].begin(), ex->mailbox[my_rank].end(), data);
        ex->posted[my_rank] = 0;
        ex->cv.notify_all();
    }
    std::vector<double> gather(double value) override
    {
        {
            std::lock_guard<std::mutex> lock(ex->mutex);
            ex->gathered[my_rank] = value;
        }
        barrier();
        std::vector<double> out;
        {
            std::lock_guard<std::mutex> lock(ex->mutex);
            out = ex->gathered;
        }
        barrier();
        return out;
    }
};

} // namespace fb
