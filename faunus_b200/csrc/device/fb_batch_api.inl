// fb_batch_trial / fb_batch_commit: host side of the windowed speculative evaluation (fb_batch.cuh).
// Included at the end of fb_api.cu (same translation unit: fb_ctx, guarded(), CUDA_CHECK, ...).

namespace {

BatchBuffers batchBuffers(fb_ctx* c, int which)
{
    BatchBuffers b{};
    b.in = c->batch.d_in[which].ptr;
    b.pold = b.in ? b.in->pold : nullptr;   // host-side address arithmetic on a device pointer
    b.idold = b.in ? b.in->idold : nullptr;
    b.table = c->batch.d_table[which].ptr;
    return b;
}

void batchAllocate(fb_ctx* c)
{
    auto& b = c->batch;
    for (int i = 0; i < 2; ++i) {
        b.d_in[i].ensure(1);
        b.d_ahead[i].ensure(1);
    }
    b.d_pair_fix.ensure(2 * kBatchMax);
    b.d_pair_redo.ensure(1);
    b.h_in.ensure(1);
}

/** (re)size the phase tables for the current k-vector cutoff */
void batchEwaldGeometry(fb_ctx* c)
{
    auto& b = c->batch;
    const int ncc = static_cast<int>(std::ceil(c->ewald.n_cutoff));
    b.geo.ncc = ncc;
    b.geo.table_stride = (ncc + 1) + 2 * (2 * ncc + 1);
    for (int i = 0; i < 3; ++i) {
        b.geo.len[i] = c->slot[0].ewald_box[i];
    }
    for (int i = 0; i < 2; ++i) {
        b.d_table[i].ensure(static_cast<size_t>(2 * kBatchMax) * b.geo.table_stride);
    }
}

int commitGrid(fb_ctx* c, bool with_ewald)
{
    if (!with_ewald) {
        return 1;
    }
    return std::max(1, std::min((c->slot[0].K + kBlock - 1) / kBlock, 2 * c->n_sm));
}

/** launch the commit kernel for the pending accepted moves (and/or to obtain Σ A_k|Q_k|²) */
void launchBatchCommit(fb_ctx* c, bool with_ewald, int* n_blocks_out)
{
    auto& b = c->batch;
    const int grid = commitGrid(c, with_ewald);
    if (with_ewald) {
        b.d_e_partials.ensure(static_cast<size_t>(grid));
    }
    CommitList list = b.has_pending ? b.pending : CommitList{};
    batchCommitKernel<<<grid, kBlock, 0, c->stream>>>(makeView(c, 0), makeView(c, 1), makeEwaldView(c, 0),
                                                      c->slot[0].kn.ptr, batchBuffers(c, b.parity), b.geo, list,
                                                      with_ewald ? 1 : 0, b.d_e_partials.ptr);
    launched(c, "batchCommitKernel");
    if (b.pending_moves.n > 0) { // group mode: mass centres of the accepted groups
        batchCommitGroupsKernel<<<1, kBatchMax, 0, c->stream>>>(makeView(c, 0), makeView(c, 1),
                                                                batchBuffers(c, b.parity), b.pending_moves);
        launched(c, "batchCommitGroupsKernel");
        b.pending_moves = CommitList{};
    }
    b.has_pending = false;
    if (n_blocks_out) {
        *n_blocks_out = with_ewald ? grid : 0;
    }
}

/** can the pair part of a window go through the device cell list? (re)computes the grid geometry */
bool cellGridFor(fb_ctx* c, CellGrid& g)
{
    auto& b = c->batch;
    if (b.force_brute || b.cell_min_particles < 0 || c->n_slots < b.cell_min_particles || c->P.any_molecular ||
        std::isinf(c->pair_cut2) || !(c->pair_cut2 > 0)) {
        return false;
    }
    const double rc = std::sqrt(c->pair_cut2);
    long n_cells = 1;
    for (int i = 0; i < 3; ++i) {
        if (!c->periodic[i]) {
            return false;
        }
        const double len = c->slot[0].box[i];
        const int n = static_cast<int>(std::floor(len / rc)); // cell edge = box / floor(box / cutoff) ≥ cutoff
        if (n < 3) {
            return false;
        }
        g.n[i] = std::min(n, 64);
        g.inv_edge[i] = g.n[i] / len;
        g.half[i] = 0.5 * len;
        n_cells *= g.n[i];
    }
    const bool same = b.cell_dims[0] == g.n[0] && b.cell_dims[1] == g.n[1] && b.cell_dims[2] == g.n[2];
    if (!same) {
        b.cells_valid = false;
        if (!b.cell_cap_forced) {
            b.cell_cap = 0;
        }
        for (int i = 0; i < 3; ++i) {
            b.cell_dims[i] = g.n[i];
        }
    }
    if (b.cell_cap == 0) {
        const double mean = static_cast<double>(c->n_slots) / static_cast<double>(n_cells);
        b.cell_cap = std::max(16, static_cast<int>(2.0 * mean) + 16);
    }
    b.d_cell_count.ensure(static_cast<size_t>(n_cells));
    b.d_cell_bucket.ensure(static_cast<size_t>(n_cells) * b.cell_cap);
    b.d_cell_overflow.ensure(1);
    g.cap = b.cell_cap;
    g.count = b.d_cell_count.ptr;
    g.bucket = b.d_cell_bucket.ptr;
    g.overflow = b.d_cell_overflow.ptr;
    return true;
}

/** build the cell list of slot 0 from scratch on `stream` (positions must be final for this window) */
void buildCellList(fb_ctx* c, const CellGrid& g, cudaStream_t stream)
{
    const long n_cells = static_cast<long>(g.n[0]) * g.n[1] * g.n[2];
    CUDA_CHECK(cudaMemsetAsync(g.count, 0, n_cells * sizeof(int), stream));
    CUDA_CHECK(cudaMemsetAsync(g.overflow, 0, sizeof(int), stream));
    cellAppendKernel<<<(c->n_slots + 255) / 256, 256, 0, stream>>>(makeView(c, 0), g);
    launched(c, "cellAppendKernel");
    cellSortKernel<<<static_cast<int>((n_cells + 127) / 128), 128, 0, stream>>>(g, static_cast<int>(n_cells));
    launched(c, "cellSortKernel");
    c->batch.cells_valid = true;
}

/** kernel launch that may carry the programmatic-dependent-launch attribute (FB_PDL; see fb_batch.cuh) */
template <class... KArgs, class... Args>
void launchChained(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool chained, Args&&... args)
{
    cudaLaunchConfig_t config = {};
    config.gridDim = grid;
    config.blockDim = block;
    config.dynamicSmemBytes = smem;
    config.stream = stream;
    cudaLaunchAttribute attribute[1];
    attribute[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attribute[0].val.programmaticStreamSerializationAllowed = 1;
    config.attrs = attribute;
    config.numAttrs = (FB_PDL && chained) ? 1 : 0;
    CUDA_CHECK(cudaLaunchKernelEx(&config, kernel, static_cast<KArgs>(args)...));
}

/**
 * Launches of one window. Main stream: prep → phase tables → k-space → k-space sums; pair stream (forked
 * after prep, joined before the copy-out): pair kernel → pair sums + cross terms. Serialised on the
 * main stream when per-kernel timing is on.
 */
template <int KIND>
void launchWindow(fb_ctx* c, const BatchBuffers& cur, const BatchBuffers& prev, const CommitList& commit,
                  const CommitList& commit_moves, int n_moves, int n_groups, int stride, bool with_ewald, bool timing,
                  bool device_commit = false, const BatchBuffers* ahead = nullptr, double fix_limit = 0.0,
                  const RunTail* tail = nullptr, bool* walked = nullptr)
{
    // tail != nullptr (runs): the window is walked on the device; *walked tells the caller whether the finish kernel
    // did that itself (windowTailKernel) or a runDecideKernel has to follow
    // ahead != nullptr (runs): the pair sums of `cur` may have been taken a window ahead, and those of `ahead` are
    // taken now, behind this window's own pair work
    // device_commit: the commit list is written on the device (a run), the host does not know whether it is empty
    auto& b = c->batch;
    const SlotView M0 = makeView(c, 0);
    const bool fork = with_ewald && !timing;
    cudaStream_t ps = fork ? b.pair_stream : c->stream;
    if (fork) {
        CUDA_CHECK(cudaEventRecord(b.ev_fork, c->stream)); // after the H2D copy of the window description
        CUDA_CHECK(cudaStreamWaitEvent(ps, b.ev_fork, 0));
    }
    const bool pair_ahead = ahead != nullptr;
    // ---- pair side
    // FP32 screening of the distance test (batchPairScreenKernel) against the widened cutoff of screeningCutoff()
    const bool screened = !pair_ahead && std::isfinite(c->pair_cut2) && c->pair_cut2 > 0 && n_groups == 0;
    const float cut2_screen = screened ? screeningCutoff(c, 0) : 0.0f;
    const int n_pair_blocks = (c->n_slots + (screened ? kScreenChunk : kPairChunk) - 1) / (screened ? kScreenChunk : kPairChunk);
    const int pair_finish_grid = (2 * stride + stride * stride + kBlock / 32 - 1) / (kBlock / 32); // one warp per output
    CellGrid grid{};
    b.cells_used = n_groups == 0 && cellGridFor(c, grid);
    // brute-force windows of single atoms: the pair kernel itself puts the previous window's accepted moves into the
    // mirrors (each block those of its own particle range) and the sums / cross terms are taken by the finish kernel
    // behind the k-space kernel: no prep launch in front of the pair kernel, no third kernel on the pair stream
    // (mass centres of accepted rigid groups — commit_moves — still go through the prep kernel)
    const bool lean = n_groups == 0 && !b.cells_used && !pair_ahead && commit_moves.n == 0;
    const bool have_commit = commit.n > 0 || commit_moves.n > 0 || device_commit;
    if (have_commit && !lean) {
        batchPrepKernel<<<1, kBatchMax, 0, ps>>>(M0, makeView(c, 1), cur, prev, pair_ahead ? b.d_pair_redo.ptr : nullptr);
        launched(c, "batchPrepKernel");
    }
    if (timing) {
        CUDA_CHECK(cudaEventRecord(b.ev[1], c->stream));
    }
    if (n_groups > 0) { // rigid-molecule moves: one block per (move, new | old), threads over the other groups
        batchPairGroupKernel<KIND><<<2 * n_groups, kBlock, 0, ps>>>(M0, c->P, cur, stride, b.d_result.ptr);
        launched(c, "batchPairGroupKernel");
        batchPairGroupFinishKernel<KIND><<<pair_finish_grid, kBlock, 0, ps>>>(M0, c->P, cur, stride, b.d_result.ptr);
        launched(c, "batchPairGroupFinishKernel");
    }
    else {
        if (b.cells_used) { // large N: 27 neighbour cells per position instead of all particles
            if (!b.cells_valid) {
                buildCellList(c, grid, ps);
            }
            else if (commit.n > 0 || device_commit) {
                cellCommitKernel<<<1, 2 * kBatchMax, 0, ps>>>(grid, cur, prev);
                launched(c, "cellCommitKernel");
            }
            batchPairCellKernel<KIND><<<2 * n_moves, kCellThreads, 0, ps>>>(M0, c->P, grid, cur, c->pair_cut2, stride,
                                                                            b.d_result.ptr);
            launched(c, "batchPairCellKernel");
        }
        else {
            b.d_pair_partials.ensure(static_cast<size_t>(n_pair_blocks) * 2 * kBatchMax);
            const dim3 pair_grid(n_pair_blocks, (2 * n_moves + kPairVariantsPerBlock - 1) / kPairVariantsPerBlock);
            const int* redo = pair_ahead ? b.d_pair_redo.ptr : nullptr;
            if (pair_ahead) { // sums taken a window ahead → corrections for what was accepted since; else: from scratch
                batchPairFixKernel<KIND><<<(2 * n_moves * 32 + kBlock - 1) / kBlock, kBlock, 0, ps>>>(
                    M0, c->P, cur, prev, fix_limit, b.d_pair_fix.ptr, b.d_pair_redo.ptr);
                launched(c, "batchPairFixKernel");
            }
            const int apply_commit = (lean && have_commit) ? 1 : 0;
            if (FB_CROSS_EARLY && lean && tail != nullptr && !timing) { // runs: the cross terms first, beside the front kernel
                windowCrossKernel<KIND><<<(stride * stride * 32 + kBlock - 1) / kBlock, kBlock, 0, ps>>>(M0, c->P, cur, stride,
                                                                                                  b.d_result.ptr);
                launched(c, "windowCrossKernel");
            }
            if (screened) { // finite cutoff: FP32 screening, FP64 evaluation of the candidates
                batchPairScreenKernel<KIND><<<pair_grid, kPairThreads, 0, ps>>>(M0, c->P, cur, c->pair_cut2, cut2_screen, stride,
                                                                               b.d_pair_partials.ptr, makeView(c, 1), prev,
                                                                               apply_commit);
            }
            else if (std::isinf(c->pair_cut2)) {
                batchPairKernel<KIND, true><<<pair_grid, kPairThreads, 0, ps>>>(M0, c->P, cur, c->pair_cut2, stride,
                                                                               b.d_pair_partials.ptr, redo, makeView(c, 1),
                                                                               prev, apply_commit);
            }
            else {
                batchPairKernel<KIND, false><<<pair_grid, kPairThreads, 0, ps>>>(M0, c->P, cur, c->pair_cut2, stride,
                                                                                b.d_pair_partials.ptr, redo, makeView(c, 1),
                                                                                prev, apply_commit);
            }
            launched(c, "batchPairKernel");
        }
        if (!lean) {
            const bool fixed = pair_ahead && !b.cells_used;
            batchPairFinishKernel<KIND><<<pair_finish_grid, kBlock, 0, ps>>>(
                M0, c->P, cur, stride, n_pair_blocks, b.d_pair_partials.ptr, b.cells_used ? 1 : 0,
                b.cells_used ? b.d_cell_overflow.ptr : nullptr, b.d_result.ptr, fixed ? b.d_pair_fix.ptr : nullptr,
                fixed ? b.d_pair_redo.ptr : nullptr);
            launched(c, "batchPairFinishKernel");
        }
    }
    if (fork) {
        CUDA_CHECK(cudaEventRecord(b.ev_join, ps));
    }
    if (pair_ahead && !b.cells_used && n_groups == 0) {
        // the pair sums of the window predicted to follow, against the positions as they are now: behind this
        // window's own pair work on the pair stream, beside its k-space kernel and its walk
        const dim3 pair_grid(n_pair_blocks, (2 * n_moves + kPairVariantsPerBlock - 1) / kPairVariantsPerBlock);
        if (std::isinf(c->pair_cut2)) {
            batchPairKernel<KIND, true><<<pair_grid, kPairThreads, 0, ps>>>(M0, c->P, *ahead, c->pair_cut2, stride,
                                                                           b.d_pair_partials.ptr);
        }
        else {
            batchPairKernel<KIND, false><<<pair_grid, kPairThreads, 0, ps>>>(M0, c->P, *ahead, c->pair_cut2, stride,
                                                                            b.d_pair_partials.ptr);
        }
        launched(c, "batchPairKernel");
    }
    if (timing) {
        CUDA_CHECK(cudaEventRecord(b.ev[2], c->stream));
    }
    // ---- k-space side: phase tables + commit of the previous window (windowFrontKernel), then the persistent kernel
    int n_rows = 0, n_e_rows = 0;
    if (with_ewald) {
        const Slot& sl = c->slot[0];
        const EwaldView E = makeEwaldView(c, 0);
        const int n_phase_blocks = frontPhaseBlocks(b.geo);
        const int n_front_blocks = sl.n_items; // one block per item
        n_e_rows = n_front_blocks;
        n_rows = std::max(1, (sl.n_sched_blocks + kKsGroups - 1) / kKsGroups); // one row of partials per block
        b.d_kq.ensure(static_cast<size_t>(sl.n_units) * kUnitSlots);
        b.d_e_partials.ensure(static_cast<size_t>(std::max(1, n_front_blocks)));
        b.d_r_partials.ensure(static_cast<size_t>(2 * c->n_sm) * kBatchMax);
        b.d_g_partials.ensure(static_cast<size_t>(2 * c->n_sm) * kBatchMax * kBatchMax);
        const bool chained = tail != nullptr && !timing; // runs: front → k-space → tail → front of the next window
        launchChained(windowFrontKernel, dim3(n_phase_blocks + n_front_blocks), dim3(kFrontThreads), 0, c->stream, chained, E,
                      sl.aks.ptr, sl.unit_info.ptr, sl.unit_map.ptr, sl.item_units.ptr, sl.item_base.ptr, n_phase_blocks, cur, prev,
                      b.geo, b.d_kq.ptr, b.d_e_partials.ptr);
        launched(c, "windowFrontKernel");
        if (timing) {
            CUDA_CHECK(cudaEventRecord(b.ev[5], c->stream));
        }
        if (!b.kspace_unit_configured) {
            CUDA_CHECK(cudaFuncSetAttribute(windowKspaceKernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            static_cast<int>(kKsGroups * sizeof(KspaceSmem))));
            b.kspace_unit_configured = true;
        }
        launchChained(windowKspaceKernel, dim3(n_rows), dim3(kKsThreads), kKsGroups * sizeof(KspaceSmem), c->stream, chained,
                      sl.unit_info.ptr, sl.unit_sa.ptr, b.d_kq.ptr, sl.unit_steps.ptr, sl.sched_first.ptr, sl.sched_units.ptr,
                      sl.n_sched_blocks, cur, b.geo, stride, b.d_r_partials.ptr, b.d_g_partials.ptr);
        launched(c, "windowKspaceKernel");
    }
    if (timing) {
        CUDA_CHECK(cudaEventRecord(b.ev[3], c->stream));
    }
    if (lean) { // k-space sums, pair sums and cross terms in one launch, behind both the k-space and the pair kernel
        if (fork) {
            CUDA_CHECK(cudaStreamWaitEvent(c->stream, b.ev_join, 0));
        }
        if (tail != nullptr) { // sums, cross terms and — by the block that finishes last — the walk
            bool* configured = b.tail_configured; // per context: function attributes belong to the device the context is on
            if (!configured[KIND]) {
                CUDA_CHECK(cudaFuncSetAttribute(windowTailKernel<KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                static_cast<int>(runDecideSmemBytes(kBatchMax))));
                configured[KIND] = true;
            }
            if (b.d_tail_ticket.ptr == nullptr) {
                b.d_tail_ticket.ensure(1);
                CUDA_CHECK(cudaMemsetAsync(b.d_tail_ticket.ptr, 0, sizeof(unsigned), c->stream));
            }
            const int cross_done = (FB_CROSS_EARLY && !timing) ? 1 : 0; // (timing: the kernels of a window are serialised and
                                                                        // attributed to pair / k-space / other as before)
#if FB_CROSS_PACKED
            const int tail_pair_warps = 2 * stride + (cross_done ? 0 : (stride * stride + 15) / 16); // sums; 16 cross terms per warp
            const int tail_pair_blocks = (tail_pair_warps + kFinishThreads / 32 - 1) / (kFinishThreads / 32);
#else
            const int tail_pair_blocks = cross_done ? (2 * stride + kFinishThreads / 32 - 1) / (kFinishThreads / 32)
                                                    : pairFinishBlocks(stride);
#endif
#if FB_TAIL_ONE_WAVE
            const int tail_grid = std::min(c->n_sm, kspaceFinishGrid(stride) + tail_pair_blocks);
#else
            const int tail_grid = kspaceFinishGrid(stride) + tail_pair_blocks;
#endif
            launchChained(windowTailKernel<KIND>, dim3(tail_grid), dim3(kFinishThreads), runDecideSmemBytes(stride), c->stream,
                          with_ewald && !timing, M0, c->P, cur, stride, with_ewald ? 1 : 0, n_rows, n_e_rows, b.d_r_partials.ptr,
                          b.d_g_partials.ptr, b.d_e_partials.ptr, n_pair_blocks, b.d_pair_partials.ptr, b.d_result.ptr,
                          b.d_tail_ticket.ptr, tail->hdr, tail->moves, tail->st, tail->next, tail->out, tail->prev_out,
                          tail->predicted, tail->ahead, cross_done);
            launched(c, "windowTailKernel");
            if (walked) {
                *walked = true;
            }
        }
        else {
            windowFinishKernel<KIND><<<kspaceFinishGrid(stride) + pairFinishBlocks(stride), kFinishThreads, 0, c->stream>>>(
                M0, c->P, cur, stride, with_ewald ? 1 : 0, n_rows, n_e_rows, b.d_r_partials.ptr, b.d_g_partials.ptr,
                b.d_e_partials.ptr, n_pair_blocks, b.d_pair_partials.ptr, 0, nullptr, b.d_result.ptr, nullptr, nullptr);
            launched(c, "windowFinishKernel");
        }
    }
    else {
        batchKspaceFinishKernel<<<kspaceFinishGrid(stride), kFinishThreads, 0, c->stream>>>(
            cur, stride, with_ewald ? 1 : 0, n_rows, n_e_rows, b.d_r_partials.ptr, b.d_g_partials.ptr, b.d_e_partials.ptr,
            b.d_result.ptr);
        launched(c, "batchKspaceFinishKernel");
        if (fork) {
            CUDA_CHECK(cudaStreamWaitEvent(c->stream, b.ev_join, 0));
        }
    }
    b.last_rec_fresh = with_ewald;
}

/** Leave windowed mode: put pending accepted moves on the device and re-align the trial slot's Q(k) */
void flushBatch(fb_ctx* c)
{
    auto& b = c->batch;
    if (b.in_flight || b.runs_in_flight > 0) { // submitted work nobody waited for: drop its results
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        CUDA_CHECK(cudaStreamSynchronize(b.run_in_stream));
        CUDA_CHECK(cudaStreamSynchronize(b.run_out_stream));
        b.in_flight = false;
        b.runs_in_flight = 0;
    }
    if (b.has_pending) {
        launchBatchCommit(c, b.pending_with_ewald, nullptr);
        b.rec_known = false;
    }
    if (b.q_dirty) {
        Slot& s0 = c->slot[0];
        Slot& s1 = c->slot[1];
        if (s1.K == s0.K && s0.K > 0) {
            CUDA_CHECK(cudaMemcpyAsync(s1.Q.ptr, s0.Q.ptr, s0.K * sizeof(double2), cudaMemcpyDeviceToDevice, c->stream));
        }
        s0.rec_valid = false;
        s1.rec_valid = false;
        b.q_dirty = false;
    }
    b.last_n = 0;
    b.rec_known = false;
    b.cells_valid = false; // whatever follows may move particles behind the cell list's back
}

} // namespace

namespace {

/** after a synchronize: device time of the window(s) between ev[0] and ev[4] (+ the per-kernel split) */
void accumulateWindowTiming(fb_ctx* c, bool timing, cudaEvent_t begin = nullptr, cudaEvent_t end = nullptr)
{
    auto& b = c->batch;
    float t04 = 0;
    CUDA_CHECK(cudaEventElapsedTime(&t04, begin ? begin : b.ev[0], end ? end : b.ev[4]));
    b.acc_total_ms += t04;
    if (timing) {
        float t01 = 0, t12 = 0, t23 = 0, t34 = 0;
        CUDA_CHECK(cudaEventElapsedTime(&t01, b.ev[0], b.ev[1]));
        CUDA_CHECK(cudaEventElapsedTime(&t12, b.ev[1], b.ev[2]));
        CUDA_CHECK(cudaEventElapsedTime(&t23, b.ev[2], b.ev[3]));
        CUDA_CHECK(cudaEventElapsedTime(&t34, b.ev[3], b.ev[4]));
        b.acc_ms[0] += t12;
        b.acc_ms[1] += t23;
        b.acc_ms[2] += t01 + t34;
        if (b.last_rec_fresh && cudaEventQuery(b.ev[5]) == cudaSuccess) { // a window with a k-space part
            float t25 = 0;
            if (cudaEventElapsedTime(&t25, b.ev[2], b.ev[5]) == cudaSuccess && t25 >= 0 && t25 <= t23) {
                b.acc_front_ms += t25;
            }
            else {
                cudaGetLastError();
            }
        }
    }
}

/** common head of the submit flavours; a run may be queued behind ONE run in flight (`chain`) */
void beginWindow(fb_ctx* c, int with_ewald, bool chain = false)
{
    checkSlot(c, 0);
    checkSlot(c, 1);
    auto& b = c->batch;
    if (b.in_flight || b.runs_in_flight > (chain ? 1 : 0)) {
        throw CudaError{"fb_batch_submit: the previous window has not been waited for"};
    }
    if (c->has_commit) { // an accepted fast-path move is still only on the host
        applyCommitKernel<<<1, 32, 0, c->stream>>>(makeView(c, 0), makeView(c, 1), c->commit);
        launched(c, "applyCommitKernel");
        c->has_commit = false;
        b.cells_valid = false;
    }
    c->trial_active = false;
    if (with_ewald) {
        if (!c->ewald_configured || c->slot[0].K <= 0) {
            throw CudaError{"fb_batch_trial: Ewald is not initialised on the accepted slot"};
        }
        if (c->ewald.policy == 2) {
            throw CudaError{"fb_batch_trial: IPBC is not supported by the windowed path"};
        }
        if (b.has_pending && !b.pending_with_ewald) {
            throw CudaError{"fb_batch_trial: windows with and without Ewald cannot be mixed"};
        }
    }
    batchAllocate(c);
}

/** launches for the window described by b.h_in (n atoms; n_groups > 0: group mode) */
void launchPreparedWindow(fb_ctx* c, int n_atoms, int n_groups, int with_ewald)
{
    auto& b = c->batch;
    const int stride = n_atoms <= 16 ? 16 : (n_atoms <= 32 ? 32 : 64);
    const bool timing = c->timing;
    CUDA_CHECK(cudaEventRecord(b.ev[0], c->stream));
    // the previous window's accepted moves (described by the buffers of parity `b.parity`): positions are
    // written by the prep kernel, their δ is added to Q(k) inside the k-space kernel
    if (with_ewald) {
        batchEwaldGeometry(c);
    }
    if (b.has_pending && b.pending_with_ewald && (!with_ewald || b.pending.n > stride)) {
        // Q(k) has to follow although this window has no k-space part / has no room for that many commits
        launchBatchCommit(c, true, nullptr);
        b.cells_valid = false; // these moves bypass the incremental cell update
    }
    const bool pending = b.has_pending;
    const CommitList commit = pending ? b.pending : CommitList{};
    const CommitList commit_moves = b.pending_moves;
    b.has_pending = false;
    b.pending_moves = CommitList{};
    const BatchBuffers prev = batchBuffers(c, b.parity);
    b.parity ^= 1;
    const BatchBuffers cur = batchBuffers(c, b.parity);
    b.h_in.ptr->commit = commit;
    b.h_in.ptr->commit_moves = commit_moves;
    b.h_in.ptr->first = 0;
    b.h_in.ptr->pair_ready = 0;
    CUDA_CHECK(cudaMemcpyAsync(cur.in, b.h_in.ptr, sizeof(BatchInput), cudaMemcpyHostToDevice, c->stream));
    const size_t n_result = batchResultDoubles(stride);
    b.d_result.ensure(batchResultDoubles(kBatchMax));
    b.h_result.ensure(batchResultDoubles(kBatchMax));
#define FB_CASE(K)                                                                                            \
    case K:                                                                                                   \
        launchWindow<K>(c, cur, prev, commit, commit_moves, n_atoms, n_groups, stride, with_ewald != 0, timing); \
        break;
    switch (c->P.kind) {
        FB_CASE(POT_COULOMB_LJ)
        FB_CASE(POT_COULOMB_WCA)
        FB_CASE(POT_PM)
        FB_CASE(POT_PMWCA)
        FB_CASE(POT_FUNCTOR)
        FB_CASE(POT_SPLINED)
    default:
        throw CudaError{"unknown potential kind"};
    }
#undef FB_CASE
    CUDA_CHECK(cudaEventRecord(b.ev[4], c->stream));
    CUDA_CHECK(cudaMemcpyAsync(b.h_result.ptr, b.d_result.ptr, n_result * sizeof(double), cudaMemcpyDeviceToHost,
                               c->stream));
    b.in_flight = true;
    b.flight_n = n_groups > 0 ? n_groups : n_atoms;
    b.flight_atoms = n_atoms;
    b.flight_groups = n_groups;
    b.flight_stride = stride;
    b.flight_with_ewald = with_ewald ? 1 : 0;
    b.flight_timing = timing;
}

} // namespace

FB_API int fb_batch_submit(fb_ctx* c, int n_moves, const fb_batch_move* moves, int with_ewald)
{
    return guarded(c, [&] {
        if (!moves || n_moves < 1 || n_moves > kBatchMax) {
            throw CudaError{"fb_batch_trial: 1..64 moves per window"};
        }
        beginWindow(c, with_ewald);
        auto& b = c->batch;
        b.last_moves.assign(moves, moves + n_moves);
        // the moved atoms: distinct, active, members of atomic groups
        BatchInput& in = *b.h_in.ptr;
        in.n = n_moves;
        in.n_groups = 0;
        in.with_ewald = with_ewald ? 1 : 0;
        for (int m = 0; m < n_moves; ++m) {
            const fb_batch_move& mv = moves[m];
            if (mv.group_index < 0 || mv.group_index >= c->n_groups) {
                throw CudaError{"fb_batch_trial: group index out of range"};
            }
            const fb_group& g = c->slot[0].groups[mv.group_index];
            if (!(c->molecule_flags[g.molid] & FB_MOL_ATOMIC)) {
                throw CudaError{"fb_batch_trial: windowed moves are single atoms of atomic groups"};
            }
            if (mv.rel_index < 0 || mv.rel_index >= g.size) {
                throw CudaError{"fb_batch_trial: relative atom index out of range"};
            }
            if (mv.atom_id < 0 || mv.atom_id >= c->P.n_types || mv.old_atom_id < 0 || mv.old_atom_id >= c->P.n_types) {
                throw CudaError{"fb_batch_trial: atom id out of range"};
            }
            in.slot[m] = g.begin + mv.rel_index;
            in.id[m] = mv.atom_id;
            in.pnew[m] = make_double4(mv.xyzq[0], mv.xyzq[1], mv.xyzq[2], mv.xyzq[3]);
            in.pold[m] = make_double4(mv.old_xyzq[0], mv.old_xyzq[1], mv.old_xyzq[2], mv.old_xyzq[3]);
            in.idold[m] = mv.old_atom_id;
            for (int a = 0; a < m; ++a) {
                if (in.slot[a] == in.slot[m]) {
                    throw CudaError{"fb_batch_trial: the moves of one window must touch distinct atoms"};
                }
            }
        }
        launchPreparedWindow(c, n_moves, 0, with_ewald);
    });
}

FB_API int fb_batch_submit_groups(fb_ctx* c, int n_moves, const fb_batch_group_move* moves, int with_ewald)
{
    return guarded(c, [&] {
        if (!moves || n_moves < 1 || n_moves > kBatchMax) {
            throw CudaError{"fb_batch_submit_groups: 1..64 moves per window"};
        }
        beginWindow(c, with_ewald);
        auto& b = c->batch;
        b.last_moves.clear(); // no cell list in group mode, hence no brute-force re-run
        BatchInput& in = *b.h_in.ptr;
        in.with_ewald = with_ewald ? 1 : 0;
        in.n_groups = n_moves;
        int n_atoms = 0;
        for (int m = 0; m < n_moves; ++m) {
            const fb_batch_group_move& mv = moves[m];
            if (mv.group_index < 0 || mv.group_index >= c->n_groups) {
                throw CudaError{"fb_batch_submit_groups: group index out of range"};
            }
            const fb_group& g = c->slot[0].groups[mv.group_index];
            if (c->molecule_flags[g.molid] & FB_MOL_ATOMIC) {
                throw CudaError{"fb_batch_submit_groups: whole-group moves are for molecular groups"};
            }
            if (mv.n_atoms != g.size || mv.n_atoms < 1 || mv.n_atoms > FB_FAST_ATOMS) {
                throw CudaError{"fb_batch_submit_groups: all (1..8) active atoms of the group must be given"};
            }
            if (n_atoms + mv.n_atoms > kBatchMax) {
                throw CudaError{"fb_batch_submit_groups: more than 64 atoms in one window"};
            }
            for (int a = 0; a < m; ++a) {
                if (in.move_group[a] == mv.group_index) {
                    throw CudaError{"fb_batch_submit_groups: the moves of one window must touch distinct groups"};
                }
            }
            in.move_first[m] = n_atoms;
            in.move_natoms[m] = mv.n_atoms;
            in.move_group[m] = mv.group_index;
            in.cm_new[m] = make_double4(mv.cm[0], mv.cm[1], mv.cm[2], 0.0);
            in.cm_old[m] = make_double4(mv.old_cm[0], mv.old_cm[1], mv.old_cm[2], 0.0);
            for (int i = 0; i < mv.n_atoms; ++i, ++n_atoms) {
                if (mv.atom_id[i] < 0 || mv.atom_id[i] >= c->P.n_types || mv.old_atom_id[i] < 0 ||
                    mv.old_atom_id[i] >= c->P.n_types) {
                    throw CudaError{"fb_batch_submit_groups: atom id out of range"};
                }
                in.slot[n_atoms] = g.begin + i;
                in.id[n_atoms] = mv.atom_id[i];
                in.idold[n_atoms] = mv.old_atom_id[i];
                in.pnew[n_atoms] = make_double4(mv.xyzq[i][0], mv.xyzq[i][1], mv.xyzq[i][2], mv.xyzq[i][3]);
                in.pold[n_atoms] = make_double4(mv.old_xyzq[i][0], mv.old_xyzq[i][1], mv.old_xyzq[i][2], mv.old_xyzq[i][3]);
            }
        }
        in.n = n_atoms;
        for (int m = 0; m < n_moves; ++m) {
            b.last_first[m] = in.move_first[m];
            b.last_natoms[m] = in.move_natoms[m];
        }
        launchPreparedWindow(c, n_atoms, n_moves, with_ewald);
    });
}

namespace {

/**
 * Launches of one window of a run: the window kernels and the walk (which also sets up the next window);
 * `setup`: the window description is not there yet (continuation after the host looked at the run).
 */
void launchRunStep(fb_ctx* c, fb_ctx::Batch::RunSlot& r, bool timing, bool setup)
{
    auto& b = c->batch;
    const int stride = r.stride;
    const bool with_ewald = r.with_ewald != 0;
    const BatchBuffers prev = batchBuffers(c, b.parity);
    b.parity ^= 1;
    const BatchBuffers cur = batchBuffers(c, b.parity);
    const BatchBuffers next = batchBuffers(c, b.parity ^ 1);
    const RunHeader* hdr = &r.d_run.ptr->header;
    const RunMove* moves = r.d_run.ptr->moves;
    RunState* st = &r.d_back.ptr->state;
    const RunOutput* prev_out = b.run[&r == &b.run[0] ? 1 : 0].d_back.ptr ? b.run[&r == &b.run[0] ? 1 : 0].d_back.ptr->out : nullptr;
    // the prediction of the window with buffer parity q lives in d_ahead[q]
    BatchInput* ahead_next = b.d_ahead[b.parity ^ 1].ptr; // of the window after this one
    BatchInput* ahead_after = b.d_ahead[b.parity].ptr;    // of the one after that (written by this step's walk)
    BatchBuffers ahead{};
    ahead.in = ahead_next;
    ahead.pold = ahead_next->pold; // host-side address arithmetic on a device pointer
    ahead.idold = ahead_next->idold;
    const bool prepair = r.h_run.ptr->header.prepair != 0;
    if (setup) {
        runSetupKernel<<<1, kBatchMax, 0, c->stream>>>(hdr, moves, st, cur.in, stride, r.d_back.ptr->out, prev_out,
                                                       ahead_next);
        launched(c, "runSetupKernel");
    }
    RunTail tail{};
    tail.hdr = hdr;
    tail.moves = moves;
    tail.st = st;
    tail.next = next.in;
    tail.out = r.d_back.ptr->out;
    tail.prev_out = prev_out;
    tail.predicted = ahead_next;
    tail.ahead = ahead_after;
    bool walked = false;
#define FB_CASE(K)                                                                                             \
    case K:                                                                                                    \
        launchWindow<K>(c, cur, prev, CommitList{}, CommitList{}, stride, 0, stride, with_ewald, timing, true,   \
                        prepair ? &ahead : nullptr, r.h_run.ptr->header.cancellation_limit, &tail, &walked);  \
        break;
    switch (c->P.kind) {
        FB_CASE(POT_COULOMB_LJ)
        FB_CASE(POT_COULOMB_WCA)
        FB_CASE(POT_PM)
        FB_CASE(POT_PMWCA)
        FB_CASE(POT_FUNCTOR)
        FB_CASE(POT_SPLINED)
    default:
        throw CudaError{"unknown potential kind"};
    }
#undef FB_CASE
    if (!walked) {
        const size_t smem = runDecideSmemBytes(stride);
        if (!b.run_decide_configured) {
            CUDA_CHECK(cudaFuncSetAttribute(runDecideKernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            static_cast<int>(runDecideSmemBytes(kBatchMax))));
            b.run_decide_configured = true;
        }
        // the walk reads the window it decides (cur) and writes the description of the next one (next == prev's
        // buffer: the window kernels of this step, the last readers of prev, are done by then)
        runDecideKernel<<<1, kDecideThreads, smem, c->stream>>>(hdr, moves, st, cur, next.in, stride, b.cells_used ? 1 : 0,
                                                               b.d_result.ptr, r.d_back.ptr->out, prev_out, ahead_next,
                                                               ahead_after);
        launched(c, "runDecideKernel");
    }
    r.steps_launched += 1;
    r.last_parity = b.parity;
}

/** everything the launches of a run's windows take their arguments from, folded into one number */
unsigned long long runGraphSignature(fb_ctx* c, const fb_ctx::Batch::RunSlot& r, int steps)
{
    auto& b = c->batch;
    unsigned long long h = 1469598103934665603ull;
    auto mix = [&](const void* data, size_t bytes) {
        const unsigned char* p = static_cast<const unsigned char*>(data);
        for (size_t i = 0; i < bytes; ++i) {
            h = (h ^ p[i]) * 1099511628211ull;
        }
    };
    auto value = [&](auto v) { mix(&v, sizeof(v)); };
    const SlotView v0 = makeView(c, 0), v1 = makeView(c, 1);
    mix(&v0, sizeof(v0));
    mix(&v1, sizeof(v1));
    const EwaldView e = makeEwaldView(c, 0);
    mix(&e, sizeof(e));
    mix(&c->P, sizeof(c->P));
    mix(&b.geo, sizeof(b.geo));
    const Slot& sl = c->slot[0];
    const auto& other = b.run[&r == &b.run[0] ? 1 : 0];
    const void* pointers[] = {b.d_in[0].ptr,       b.d_in[1].ptr,       b.d_ahead[0].ptr,    b.d_ahead[1].ptr,   b.d_table[0].ptr,
                              b.d_table[1].ptr,    b.d_pair_partials.ptr, b.d_r_partials.ptr, b.d_g_partials.ptr, b.d_e_partials.ptr,
                              b.d_result.ptr,      b.d_kq.ptr,          b.d_tail_ticket.ptr, r.d_run.ptr,        r.d_back.ptr,
                              other.d_back.ptr,    sl.aks.ptr,          sl.unit_info.ptr,    sl.unit_map.ptr,    sl.unit_sa.ptr,
                              sl.unit_steps.ptr,   sl.sched_first.ptr,  sl.sched_units.ptr,  sl.item_units.ptr,  sl.item_base.ptr,
                              b.d_pair_fix.ptr,    b.d_pair_redo.ptr};
    mix(pointers, sizeof(pointers));
    value(sl.n_units);
    value(sl.n_items);
    value(sl.n_sched_blocks);
    value(c->n_slots);
    value(c->n_sm);
    value(c->pair_cut2);
    value(screeningCutoff(c, 0));
    value(b.parity);
    value(steps);
    value(r.stride);
    value(r.with_ewald);
    value(r.h_run.ptr->header.cancellation_limit);
    return h;
}

/**
 * The windows of a run as ONE graph launch. A run is 4 kernels per window on two streams (fork / join events), 8
 * windows at S1: between dependent kernels the stream scheduler leaves ≈ 3.7 µs (11 µs per 76 µs window). The same
 * sequence — same buffers, same parity, same number of windows — comes up again and again, so the second time it is
 * captured (the first time went through plain launches: allocations and function attributes are settled) and from
 * then on replayed. Everything the kernels read that changes from run to run lives in device memory. Returns false
 * when the plain path has to run (first sight, cell list, pair sums ahead, a run continued after the host looked).
 */
bool launchRunGraph(fb_ctx* c, fb_ctx::Batch::RunSlot& r, int steps, bool continuation, int min_seen = 1)
{
    auto& b = c->batch;
    static const bool switched_off = std::getenv("FAUNUS_B200_NO_GRAPHS") != nullptr; // experiments: plain launches
    if (switched_off || !b.run_graphs_enabled || continuation || steps < 1 || r.h_run.ptr->header.prepair != 0 || r.with_ewald == 0) {
        return false;
    }
    CellGrid grid{};
    if (cellGridFor(c, grid)) {
        return false;
    }
    if (b.run_graphs.size() > 64) { // a box that keeps changing (NPT) leaves keys behind that never come back
        CUDA_CHECK(cudaStreamSynchronize(c->stream)); // (rare) nothing of them is in flight when they go
        for (auto& [key, g] : b.run_graphs) {
            if (g.exec) {
                cudaGraphExecDestroy(g.exec);
            }
        }
        b.run_graphs.clear();
    }
    auto& entry = b.run_graphs[runGraphSignature(c, r, steps)];
    if (entry.exec == nullptr) {
        if (entry.seen++ < min_seen) {
            return false;
        }
        const long launches_before = c->launches;
        CUDA_CHECK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed));
        cudaGraph_t graph = nullptr;
        try {
            for (int s = 0; s < steps; ++s) {
                launchRunStep(c, r, false, false);
            }
        }
        catch (...) {
            cudaStreamEndCapture(c->stream, &graph);
            if (graph) {
                cudaGraphDestroy(graph);
            }
            throw;
        }
        CUDA_CHECK(cudaStreamEndCapture(c->stream, &graph));
        CUDA_CHECK(cudaGraphInstantiate(&entry.exec, graph, 0));
        CUDA_CHECK(cudaGraphDestroy(graph));
        entry.launches = c->launches - launches_before;
        CUDA_CHECK(cudaGraphLaunch(entry.exec, c->stream)); // the capture pass already advanced the host-side state
        b.run_graph_replays += 1;
        return true;
    }
    CUDA_CHECK(cudaGraphLaunch(entry.exec, c->stream));
    // what launchRunStep / launchWindow leave behind on the host
    b.parity ^= (steps & 1);
    b.cells_used = false;
    b.last_rec_fresh = r.with_ewald != 0;
    r.steps_launched += steps;
    r.last_parity = b.parity;
    c->launches += entry.launches;
    b.run_graph_replays += 1;
    return true;
}

/** queue `steps` windows of the run and the read-back of where it stands */
void launchRunSteps(fb_ctx* c, fb_ctx::Batch::RunSlot& r, int steps, bool continuation)
{
    auto& b = c->batch;
    if (c->timing) { // per-kernel events: one window at a time, kernels serialised
        for (int s = 0; s < steps; ++s) {
            CUDA_CHECK(cudaEventRecord(b.ev[0], c->stream));
            launchRunStep(c, r, true, continuation && s == 0);
            CUDA_CHECK(cudaEventRecord(b.ev[4], c->stream));
            CUDA_CHECK(cudaStreamSynchronize(c->stream));
            accumulateWindowTiming(c, true);
        }
    }
    else {
        CUDA_CHECK(cudaEventRecord(r.ev_begin, c->stream));
        // Two granularities. One graph per WINDOW: four keys (run slot × buffer parity), all captured within the first few
        // runs. One graph per RUN: a few µs better per window (measured 8.7–8.9e5 against 8.46e5 moves/s at S1), but the
        // number of windows of a run varies (7 … 11 at S1) and every capture + instantiation costs ≈ 3 ms — so a run's
        // sequence is only captured once it has come up kRunGraphMinSeen times; the rare ones go window by window.
        constexpr int kRunGraphMinSeen = 4;
        if (steps < 2 || !launchRunGraph(c, r, steps, continuation, kRunGraphMinSeen)) {
            for (int s = 0; s < steps; ++s) {
                const bool setup = continuation && s == 0;
                if (!launchRunGraph(c, r, 1, setup)) {
                    launchRunStep(c, r, false, setup);
                }
            }
        }
        CUDA_CHECK(cudaEventRecord(r.ev_end, c->stream));
    }
    // state + overflow flag + the decisions
    if (b.cells_used) { // (read by fb_run_wait with the cell list only; a copy node costs a few µs between two runs)
        CUDA_CHECK(cudaMemcpyAsync(&r.d_back.ptr->overflow, b.d_result.ptr + 2, sizeof(double), cudaMemcpyDeviceToDevice,
                                   c->stream));
    }
    // … on the read-back stream: the next run's kernels need not wait for the copy (they only READ this run's state and
    // decisions; its slot is written again by the run after the next, which the host submits after this wait)
    const size_t bytes = offsetof(fb_ctx::Batch::RunBack, out) + sizeof(RunOutput) * static_cast<size_t>(r.n);
    CUDA_CHECK(cudaEventRecord(r.ev_tail, c->stream));
    CUDA_CHECK(cudaStreamWaitEvent(b.run_out_stream, r.ev_tail, 0));
    CUDA_CHECK(cudaMemcpyAsync(r.h_back.ptr, r.d_back.ptr, bytes, cudaMemcpyDeviceToHost, b.run_out_stream));
    CUDA_CHECK(cudaEventRecord(r.ev_done, b.run_out_stream));
}

/** first launches of a run: its start state (from the host's pending list, or chained to the run before it) */
void launchRun(fb_ctx* c, fb_ctx::Batch::RunSlot& r, const fb_ctx::Batch::RunSlot* behind, const CommitList& pending)
{
    auto& b = c->batch;
    // the first window goes into the buffer the first step will call `cur`
    BatchInput* first = batchBuffers(c, b.parity ^ 1).in;
    auto& other = b.run[&r == &b.run[0] ? 1 : 0];
    const RunOutput* prev_out = other.d_back.ptr ? other.d_back.ptr->out : nullptr;
    if (behind != nullptr) {
        runChainKernel<<<1, kBatchMax, 0, c->stream>>>(&behind->d_run.ptr->header, &behind->d_back.ptr->state,
                                                       &r.d_run.ptr->header, r.d_run.ptr->moves, &r.d_back.ptr->state,
                                                       first, r.stride, r.d_back.ptr->out, prev_out,
                                                       b.d_ahead[b.parity].ptr);
        launched(c, "runChainKernel");
    }
    else {
        runInitKernel<<<1, kBatchMax, 0, c->stream>>>(&r.d_run.ptr->header, r.d_run.ptr->moves, &r.d_back.ptr->state,
                                                      pending, first, r.stride, r.d_back.ptr->out, prev_out,
                                                      b.d_ahead[b.parity].ptr);
        launched(c, "runInitKernel");
    }
    CUDA_CHECK(cudaEventRecord(r.ev_started, c->stream));
    r.chained = behind != nullptr;
    r.steps_launched = 0;
    // windows the run needs (barring cancellations): a window ends before a proposal that depends on one of its moves
    int steps = 0;
    for (int cursor = 0; cursor < r.n; ++steps) {
        int len = std::min(r.stride, r.n - cursor);
        for (int t = 1; t < len; ++t) {
            const int dep = r.h_run.ptr->moves[cursor + t].dep;
            if (dep >= cursor && !(dep & kRunDepPrevious)) {
                len = t;
            }
        }
        cursor += len;
    }
    launchRunSteps(c, r, steps, false);
}

} // namespace

FB_API int fb_run_submit(fb_ctx* c, int n_moves, const fb_run_move* moves, int with_ewald, const fb_run_config* config)
{
    return guarded(c, [&] {
        if (!moves || !config || n_moves < 1 || n_moves > kRunMax) {
            throw CudaError{"fb_run_submit: 1..1024 moves per run"};
        }
        auto& b = c->batch;
        const bool chain = b.runs_in_flight == 1;
        beginWindow(c, with_ewald, chain);
        auto& r = b.run[(b.run_head + b.runs_in_flight) % 2];
        const fb_ctx::Batch::RunSlot* behind = chain ? &b.run[b.run_head] : nullptr;
        if (chain && behind->with_ewald != (with_ewald ? 1 : 0)) {
            throw CudaError{"fb_run_submit: runs with and without Ewald cannot be queued behind each other"};
        }
        r.h_run.ensure(1);
        r.d_run.ensure(1);
        r.h_back.ensure(1);
        r.d_back.ensure(1);
        if (static_cast<int>(b.run_stamp.size()) != c->n_slots) {
            b.run_stamp.assign(static_cast<size_t>(c->n_slots), 0);
            b.run_last.assign(static_cast<size_t>(c->n_slots), 0);
            b.run_id = 0;
        }
        b.run_id += 1;
        r.id = b.run_id;
        RunHeader& h = r.h_run.ptr->header;
        h.n_moves = n_moves;
        h.with_ewald = with_ewald ? 1 : 0;
        h.max_energy = config->max_energy;
        h.cancellation_limit = config->cancellation_limit;
        h.pad = 0;
        {
            CellGrid grid{};
            h.prepair = (b.prepair && !c->timing && !cellGridFor(c, grid)) ? 1 : 0;
        }
        h.rec_prefactor = 0.0;
        if (with_ewald) {
            batchEwaldGeometry(c);
            const double pi = 3.141592653589793238462643383279502884;
            const Slot& sl = c->slot[0];
            h.rec_prefactor = 2 * pi * c->ewald.bjerrum_length / (sl.ewald_box[0] * sl.ewald_box[1] * sl.ewald_box[2]);
        }
        for (int m = 0; m < n_moves; ++m) {
            const fb_batch_move& mv = moves[m].move;
            if (mv.group_index < 0 || mv.group_index >= c->n_groups) {
                throw CudaError{"fb_run_submit: group index out of range"};
            }
            const fb_group& g = c->slot[0].groups[mv.group_index];
            if (!(c->molecule_flags[g.molid] & FB_MOL_ATOMIC)) {
                throw CudaError{"fb_run_submit: the moves of a run are single atoms of atomic groups"};
            }
            if (mv.rel_index < 0 || mv.rel_index >= g.size) {
                throw CudaError{"fb_run_submit: relative atom index out of range"};
            }
            if (mv.atom_id < 0 || mv.atom_id >= c->P.n_types || mv.old_atom_id < 0 || mv.old_atom_id >= c->P.n_types) {
                throw CudaError{"fb_run_submit: atom id out of range"};
            }
            const int slot = g.begin + mv.rel_index;
            const int dep = moves[m].depends_on;
            RunMove& rm = r.h_run.ptr->moves[m];
            if (dep < 0) { // an ordinary proposal: nobody else in flight touches its atom
                if (b.run_stamp[slot] == r.id || (chain && b.run_stamp[slot] == behind->id)) {
                    throw CudaError{"fb_run_submit: the moves of a run (and of the run it is queued behind) must touch "
                                    "distinct atoms unless they name the move they depend on"};
                }
            }
            else { // it starts where move `dep` leaves the atom: same atom, earlier, and nobody in between
                const bool previous = (dep & FB_RUN_DEP_PREVIOUS) != 0;
                const int index = dep & (FB_RUN_DEP_PREVIOUS - 1);
                const RunMove* target = nullptr;
                if (previous) {
                    target = (chain && index < behind->n) ? &behind->h_run.ptr->moves[index] : nullptr;
                }
                else {
                    target = index < m ? &r.h_run.ptr->moves[index] : nullptr;
                }
                if (target == nullptr || target->slot != slot || b.run_stamp[slot] != (previous ? behind->id : r.id) ||
                    b.run_last[slot] != index) {
                    throw CudaError{"fb_run_submit: depends_on must name the latest earlier move on the same atom"};
                }
                const fb_batch_move& av = moves[m].alt;
                if (av.atom_id < 0 || av.atom_id >= c->P.n_types || av.old_atom_id < 0 || av.old_atom_id >= c->P.n_types) {
                    throw CudaError{"fb_run_submit: atom id out of range"};
                }
                rm.pnew_alt = make_double4(av.xyzq[0], av.xyzq[1], av.xyzq[2], av.xyzq[3]);
                rm.pold_alt = make_double4(av.old_xyzq[0], av.old_xyzq[1], av.old_xyzq[2], av.old_xyzq[3]);
            }
            b.run_stamp[slot] = r.id;
            b.run_last[slot] = m;
            rm.slot = slot;
            rm.id = mv.atom_id;
            rm.idold = mv.old_atom_id;
            rm.flags = moves[m].flags;
            rm.dep = dep;
            rm.pad = 0;
            rm.pnew = make_double4(mv.xyzq[0], mv.xyzq[1], mv.xyzq[2], mv.xyzq[3]);
            rm.pold = make_double4(mv.old_xyzq[0], mv.old_xyzq[1], mv.old_xyzq[2], mv.old_xyzq[3]);
            rm.uniform = moves[m].uniform;
            rm.host_new = moves[m].host_new;
            rm.host_old = moves[m].host_old;
            rm.host_new_alt = moves[m].alt_host_new;
            rm.host_old_alt = moves[m].alt_host_old;
        }
        r.n = n_moves;
        r.with_ewald = with_ewald ? 1 : 0;
        // behind another run the size of the commit list it leaves is not known here: full windows
        r.stride = (chain || n_moves > 32) ? 64 : (n_moves <= 16 ? 16 : 32);
        // accepted moves of an earlier window / run that are not yet on the device
        if (!chain && b.has_pending &&
            (b.pending_moves.n > 0 || (b.pending_with_ewald && (!with_ewald || b.pending.n > r.stride)))) {
            // rigid-molecule moves (mass centres) / Q(k) has to follow although this run has no k-space part / its
            // windows have no room for that many commits: apply them now
            launchBatchCommit(c, b.pending_with_ewald, nullptr);
            b.cells_valid = false;
        }
        const CommitList pending = (!chain && b.has_pending) ? b.pending : CommitList{};
        b.has_pending = false;
        b.d_result.ensure(batchResultDoubles(kBatchMax));
        b.h_result.ensure(batchResultDoubles(kBatchMax));
        // the input goes up beside the kernels of the run in flight. That run's chain kernel read the header of the run
        // that had this slot before: the upload waits for it (long through in practice)
        const size_t bytes = offsetof(fb_ctx::Batch::RunBlock, moves) + sizeof(RunMove) * static_cast<size_t>(n_moves);
        if (behind != nullptr) {
            CUDA_CHECK(cudaStreamWaitEvent(b.run_in_stream, behind->ev_started, 0));
        }
        CUDA_CHECK(cudaMemcpyAsync(r.d_run.ptr, r.h_run.ptr, bytes, cudaMemcpyHostToDevice, b.run_in_stream));
        CUDA_CHECK(cudaEventRecord(r.ev_in, b.run_in_stream));
        CUDA_CHECK(cudaStreamWaitEvent(c->stream, r.ev_in, 0));
        b.runs_in_flight += 1;
        launchRun(c, r, behind, pending);
    });
}

FB_API int fb_run_wait(fb_ctx* c, fb_run_result* out)
{
    return guarded(c, [&] {
        auto& b = c->batch;
        if (!out || b.runs_in_flight < 1) {
            throw CudaError{"fb_run_wait: no submitted run"};
        }
        auto& r = b.run[b.run_head];
        fb_ctx::Batch::RunSlot* successor = b.runs_in_flight == 2 ? &b.run[b.run_head ^ 1] : nullptr;
        bool continued = false;
        for (;;) {
            CUDA_CHECK(cudaEventSynchronize(r.ev_done));
            if (!c->timing) {
                accumulateWindowTiming(c, false, r.ev_begin, r.ev_end);
            }
            const RunState& st = r.h_back.ptr->state;
            if (st.cursor >= r.n) {
                break;
            }
            if (r.steps_launched > 4 * kRunMax) {
                throw CudaError{"fb_run_wait: the run makes no progress"};
            }
            // a window stopped early (cancellation) or a cell bucket ran full: more windows. A run queued behind
            // this one has halted on the device; its (empty) windows are in the stream before what follows.
            if (b.cells_used && r.h_back.ptr->overflow != 0.0) {
                b.cells_valid = false;
                b.cell_cap *= 2;
            }
            if (!continued && successor != nullptr) {
                b.parity = r.last_parity; // the window buffers as this run left them (its last window is `prev`)
            }
            continued = true;
            const int left = r.n - st.cursor;
            launchRunSteps(c, r, (left + r.stride - 1) / r.stride, true);
        }
        if (b.gap_stats && successor != nullptr && !continued && !c->timing) {
            // experiments (FAUNUS_B200_GAP_STATS): device time between the last window of this run and the first window
            // of the run queued behind it (chain kernel, event records, launch latency)
            float gap_ms = 0.0f;
            if (cudaEventSynchronize(successor->ev_begin) == cudaSuccess &&
                cudaEventElapsedTime(&gap_ms, r.ev_end, successor->ev_begin) == cudaSuccess) {
                b.gap_ms_total += gap_ms;
                b.gap_count += 1;
            }
            cudaGetLastError();
        }
        const RunState& st = r.h_back.ptr->state;
        if (successor != nullptr && continued) { // launch the halted run again, from the state this one leaves
            launchRun(c, *successor, nullptr, st.commit);
        }
        b.runs_in_flight -= 1;
        b.run_head ^= 1;
        b.windows += st.steps;
        b.moves += r.n;
        b.run_steps += st.steps;
        b.run_moves += r.n;
        b.run_count += 1;
        b.round_trips += 1;
        b.run_accepted.resize(static_cast<size_t>(r.n));
        b.run_u_new.resize(static_cast<size_t>(r.n));
        b.run_u_old.resize(static_cast<size_t>(r.n));
        bool any = false;
        for (int m = 0; m < r.n; ++m) {
            const RunOutput& o = r.h_back.ptr->out[m];
            b.run_accepted[m] = static_cast<unsigned char>(o.accepted != 0);
            b.run_u_new[m] = o.u_new;
            b.run_u_old[m] = o.u_old;
            any = any || o.accepted != 0;
        }
        // the accepted moves of the last window are still to be applied: by the run queued behind this one (on the
        // device), else by the next window, run or flush
        if (successor == nullptr) {
            b.pending = st.commit;
            b.has_pending = st.commit.n > 0;
            b.pending_moves = CommitList{};
            b.pending_with_ewald = r.with_ewald != 0;
        }
        b.last_n = 0;
        b.last_groups = 0;
        if (any && r.with_ewald) {
            b.q_dirty = true;
            c->slot[0].rec_valid = false;
            c->slot[1].rec_valid = false;
        }
        b.rec_known = false;
        out->n_moves = r.n;
        out->n_windows = st.steps;
        out->n_rounds = st.rounds;
        b.run_rounds += st.rounds;
        out->accepted = b.run_accepted.data();
        out->u_new = b.run_u_new.data();
        out->u_old = b.run_u_old.data();
    });
}

FB_API int fb_batch_wait(fb_ctx* c, fb_batch_result* out)
{
    return guarded(c, [&] {
        auto& b = c->batch;
        if (!out || !b.in_flight) {
            throw CudaError{"fb_batch_wait: no submitted window"};
        }
        const int n_moves = b.flight_n;
        const int stride = b.flight_stride;
        const int with_ewald = b.flight_with_ewald;
        const bool timing = b.flight_timing;
        b.in_flight = false;
        CUDA_CHECK(cudaStreamSynchronize(c->stream));
        accumulateWindowTiming(c, timing);
        b.windows += 1;
        b.round_trips += 1;
        b.moves += n_moves;
        const double* r = b.h_result.ptr;
        if (b.cells_used && b.flight_groups == 0 && r[2] != 0.0) { // a bucket ran full: this window's pair sums are incomplete
            b.cells_valid = false;
            b.cell_cap *= 2;
            b.force_brute = true; // the commits of this window are already on the device: plain re-evaluation
            const std::vector<fb_batch_move> again = b.last_moves;
            const int rc = fb_batch_submit(c, n_moves, again.data(), with_ewald);
            b.force_brute = false;
            if (rc != FB_OK) {
                throw CudaError{c->last_error};
            }
            const int rc2 = fb_batch_wait(c, out);
            if (rc2 != FB_OK) {
                throw CudaError{c->last_error};
            }
            return;
        }
        if (with_ewald && b.last_rec_fresh) {
            b.rec_sum = r[0];
            b.rec_known = true;
        }
        b.last_n = n_moves;
        b.last_groups = b.flight_groups;
        b.last_with_ewald = with_ewald ? 1 : 0;
        const size_t S = static_cast<size_t>(stride);
        out->n_atoms = b.flight_atoms;
        out->n_moves = n_moves;
        out->stride = stride;
        out->u_new = r + 8;
        out->u_old = r + 8 + S;
        out->rec_delta = r + 8 + 2 * S;
        out->cross_new = r + 8 + 3 * S;
        out->cross_old = r + 8 + 3 * S + S * S;
        out->cross_max = r + 8 + 3 * S + 2 * S * S;
        out->rec_cross = r + 8 + 3 * S + 3 * S * S;
        out->rec_start = 0.0;
        out->rec_prefactor = 0.0;
        if (with_ewald) {
            const double pi = 3.141592653589793238462643383279502884;
            const Slot& sl = c->slot[0];
            out->rec_prefactor = 2 * pi * c->ewald.bjerrum_length / (sl.ewald_box[0] * sl.ewald_box[1] * sl.ewald_box[2]);
            out->rec_start = b.rec_sum;
        }
    });
}

FB_API int fb_batch_trial(fb_ctx* c, int n_moves, const fb_batch_move* moves, int with_ewald, fb_batch_result* out)
{
    const int rc = fb_batch_submit(c, n_moves, moves, with_ewald);
    return rc != FB_OK ? rc : fb_batch_wait(c, out);
}

FB_API int fb_batch_commit(fb_ctx* c, int n_decided, const unsigned char* accepted)
{
    return guarded(c, [&] {
        auto& b = c->batch;
        if (b.last_n <= 0) {
            throw CudaError{"fb_batch_commit: no evaluated window"};
        }
        if (n_decided < 0 || n_decided > b.last_n || (n_decided > 0 && !accepted)) {
            throw CudaError{"fb_batch_commit: bad arguments"};
        }
        if (b.has_pending) {
            throw CudaError{"fb_batch_commit: the window was already committed"};
        }
        CommitList list{};
        CommitList moves{};
        for (int m = 0; m < n_decided; ++m) {
            if (!accepted[m]) {
                continue;
            }
            if (b.last_groups > 0) { // group mode: every atom of the accepted group, and its mass centre
                moves.index[moves.n++] = m;
                for (int i = 0; i < b.last_natoms[m]; ++i) {
                    list.index[list.n++] = b.last_first[m] + i;
                }
            }
            else {
                list.index[list.n++] = m;
            }
        }
        b.last_n = 0;
        if (list.n == 0) {
            return;
        }
        b.pending = list;
        b.pending_moves = moves;
        b.has_pending = true;
        b.pending_with_ewald = b.last_with_ewald != 0;
        if (b.pending_with_ewald) {
            b.q_dirty = true;
            b.rec_known = false;
            c->slot[0].rec_valid = false;
            c->slot[1].rec_valid = false;
        }
    });
}

FB_API int fb_configure_cells(fb_ctx* c, int min_particles)
{
    if (!c) {
        return FB_ERR_INVALID;
    }
    c->batch.cell_min_particles = min_particles;
    c->batch.cells_valid = false;
    return FB_OK;
}

/** test hook: start from this bucket capacity (the library doubles it whenever a bucket runs full) */
FB_API int fb_debug_set_cell_capacity(fb_ctx* c, int capacity)
{
    if (!c || capacity < 1) {
        return FB_ERR_INVALID;
    }
    c->batch.cell_cap = capacity;
    c->batch.cell_cap_forced = true;
    c->batch.cells_valid = false;
    return FB_OK;
}

FB_API int fb_configure_runs(fb_ctx* c, int pair_sums_ahead)
{
    if (!c) {
        return FB_ERR_INVALID;
    }
    c->batch.prepair = (pair_sums_ahead & 1) != 0;
    c->batch.run_graphs_enabled = (pair_sums_ahead & 2) == 0;
    return FB_OK;
}

FB_API int fb_get_run_stats(const fb_ctx* c, double out[4])
{
    if (!c || !out) {
        return FB_ERR_INVALID;
    }
    out[0] = c->batch.run_count;
    out[1] = c->batch.run_steps;
    out[2] = c->batch.run_rounds;
    out[3] = c->batch.run_moves;
    return FB_OK;
}

FB_API int fb_get_kspace_timing(const fb_ctx* c, double out[2])
{
    if (!c || !out) {
        return FB_ERR_INVALID;
    }
    out[0] = c->batch.acc_front_ms;
    out[1] = c->batch.acc_ms[1] - c->batch.acc_front_ms;
    return FB_OK;
}

FB_API int fb_get_batch_timing(const fb_ctx* c, double out[8])
{
    if (!c || !out) {
        return FB_ERR_INVALID;
    }
    out[0] = c->batch.acc_ms[0];
    out[1] = c->batch.acc_ms[1];
    out[2] = c->batch.acc_ms[2];
    out[3] = c->batch.windows;
    out[4] = c->batch.moves;
    out[5] = c->batch.acc_total_ms;
    out[6] = c->batch.round_trips;
    out[7] = c->batch.run_count;
    return FB_OK;
}
