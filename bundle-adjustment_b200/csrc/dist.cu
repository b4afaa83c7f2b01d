// dist.cu -- multi-GPU plumbing: NCCL (bound at run time from the libnccl the process already has, i.e. the one
// torch.distributed uses) and the panel broadcast of the distributed Cholesky (dense_driver.hpp::potrf_distributed).
//
// One process per GPU.  Collectives used by the path:
//   * ncclAllReduce(sum, f64) of the shared normal-equation pieces after the image-sharded assembly
//     (per-point partials, per-camera sums, the EO row strip of N, n, Omega);
//   * ncclBroadcast of every factored block-column panel (NVLink 5 / NVSwitch: every peer one hop away);
//   * nothing for the inverse: with the factor replicated every GPU inverts its own column tiles.
#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <stdexcept>

#include "common.h"
#include "dist.h"

namespace jaicov {

namespace {
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId *) = nullptr;
    int (*CommInitRank)(void **, int, NcclUniqueId, int) = nullptr;
    int (*CommInitAll)(void **, int, const int *) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;

template <class F>
void bind(F &f, const char *name) {
    f = reinterpret_cast<F>(dlsym(g_nccl.lib, name));
    if (!f) throw std::runtime_error(std::string("NCCL symbol missing: ") + name);
}

void load_nccl() {
    if (g_nccl.lib) return;
    // prefer the copy that is already mapped into the process (torch's), so that only one NCCL exists
    void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) throw std::runtime_error("libnccl.so.2 not found (multi-GPU needs NCCL)");
    g_nccl.lib = lib;
    bind(g_nccl.GetUniqueId, "ncclGetUniqueId");
    bind(g_nccl.CommInitRank, "ncclCommInitRank");
    bind(g_nccl.CommInitAll, "ncclCommInitAll");
    bind(g_nccl.Broadcast, "ncclBroadcast");
    bind(g_nccl.AllReduce, "ncclAllReduce");
    bind(g_nccl.CommDestroy, "ncclCommDestroy");
    bind(g_nccl.GetErrorString, "ncclGetErrorString");
}

void nccl_check(int rc, const char *what) {
    if (rc != 0) throw std::runtime_error(std::string("NCCL error in ") + what + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?"));
}
constexpr int kNcclFloat64 = 8, kNcclSum = 0;
}  // namespace

void nccl_unique_id(NcclUniqueId *out) {
    load_nccl();
    nccl_check(g_nccl.GetUniqueId(out), "ncclGetUniqueId");
}

// one communicator per device of ONE process (the single-process multi-GPU handle): comms[i] belongs to devices[i]
void nccl_comm_init_all(void **comms, int n, const int *devices) {
    load_nccl();
    nccl_check(g_nccl.CommInitAll(comms, n, devices), "ncclCommInitAll");
}

void DistContext::init(int rank_, int world_, const NcclUniqueId &id) {
    load_nccl();
    void *c = nullptr;
    nccl_check(g_nccl.CommInitRank(&c, world_, id, rank_), "ncclCommInitRank");
    adopt(rank_, world_, c);
}

// streams and events on the CURRENT device (the one the communicator belongs to)
void DistContext::adopt(int rank_, int world_, void *comm_) {
    rank = rank_;
    world = world_;
    comm = comm_;
    int lo = 0, hi = 0;
    JCHECK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    JCHECK(cudaStreamCreateWithPriority(&net, cudaStreamNonBlocking, hi));
    for (int b = 0; b < 2; b++) {
        JCHECK(cudaEventCreateWithFlags(&ev_ready[b], cudaEventDisableTiming));
        JCHECK(cudaEventCreateWithFlags(&ev_bcast[b], cudaEventDisableTiming));
        JCHECK(cudaEventCreateWithFlags(&ev_unpacked[b], cudaEventDisableTiming));
    }
    JCHECK(cudaEventCreateWithFlags(&ev_tmp, cudaEventDisableTiming));
}

void DistContext::destroy() {
    if (comm) { g_nccl.CommDestroy(comm); comm = nullptr; }
    if (net) { cudaStreamDestroy(net); net = nullptr; }
    for (int b = 0; b < 2; b++) {
        if (ev_ready[b]) cudaEventDestroy(ev_ready[b]);
        if (ev_bcast[b]) cudaEventDestroy(ev_bcast[b]);
        if (ev_unpacked[b]) cudaEventDestroy(ev_unpacked[b]);
        if (stage[b]) cudaFree(stage[b]);
        ev_ready[b] = ev_bcast[b] = ev_unpacked[b] = nullptr;
        stage[b] = nullptr;
    }
    if (ev_tmp) { cudaEventDestroy(ev_tmp); ev_tmp = nullptr; }
}

// in-place sum over ranks, enqueued on the compute stream
void DistContext::allreduce_sum(double *buf, size_t count, cudaStream_t s) {
    if (world <= 1 || count == 0) return;
    nccl_check(g_nccl.AllReduce(buf, buf, count, kNcclFloat64, kNcclSum, comm, s), "ncclAllReduce");
}

void launch_copy2d(double *dst, int64_t ldd, const double *src, int64_t lds, int64_t rows, int64_t cols, cudaStream_t s);

void PanelComm::ensure_stage(size_t elems) {
    if (ctx->stage_elems >= elems) return;
    for (int b = 0; b < 2; b++) {
        if (ctx->stage[b]) cudaFree(ctx->stage[b]);
        JCHECK(cudaMalloc(&ctx->stage[b], elems * sizeof(double)));
    }
    ctx->stage_elems = elems;
    used[0] = used[1] = false;
}

// owner: the panel starting at block column p has just been factored on the compute stream -> pack it
void PanelComm::panel_ready(int p) {
    const int b = p & 1;
    const int64_t c0 = (int64_t)p * pw * kBlk;
    const int64_t cols = std::min<int64_t>((int64_t)pw * kBlk, np - c0), rows = np - c0;
    if (used[b]) JCHECK(cudaStreamWaitEvent(compute, ctx->ev_bcast[b], 0));   // previous broadcast out of this buffer is done
    launch_copy2d(ctx->stage[b], cols, M + c0 * ld + c0, ld, rows, cols, compute);
    launch_copy2d(ctx->stage[b] + rows * cols, kBlk, Dinv + c0 * kBlk, kBlk, cols, kBlk, compute);
    JCHECK(cudaEventRecord(ctx->ev_ready[b], compute));
}

void PanelComm::bcast_panel(int k, int64_t row0, int64_t rows, int64_t col0, int64_t cols, int root) {
    const int b = k & 1;
    const size_t count = (size_t)rows * cols + (size_t)cols * kBlk;
    if (ctx->rank == root) {
        JCHECK(cudaStreamWaitEvent(ctx->net, ctx->ev_ready[b], 0));
        nccl_check(g_nccl.Broadcast(ctx->stage[b], ctx->stage[b], count, kNcclFloat64, root, ctx->comm, ctx->net), "ncclBroadcast");
        JCHECK(cudaEventRecord(ctx->ev_bcast[b], ctx->net));
    } else {
        if (used[b]) JCHECK(cudaStreamWaitEvent(ctx->net, ctx->ev_unpacked[b], 0));   // buffer content (panel k-2) consumed
        nccl_check(g_nccl.Broadcast(ctx->stage[b], ctx->stage[b], count, kNcclFloat64, root, ctx->comm, ctx->net), "ncclBroadcast");
        JCHECK(cudaEventRecord(ctx->ev_bcast[b], ctx->net));
        JCHECK(cudaStreamWaitEvent(compute, ctx->ev_bcast[b], 0));
        launch_copy2d(M + row0 * ld + col0, ld, ctx->stage[b], cols, rows, cols, compute);
        launch_copy2d(Dinv + col0 * kBlk, kBlk, ctx->stage[b] + rows * cols, kBlk, cols, kBlk, compute);
        JCHECK(cudaEventRecord(ctx->ev_unpacked[b], compute));
    }
    used[b] = true;
}

// ---- owner-only storage: panels are consumed in the staging slots -------------------------------------------------------------------
void StreamPanelComm::ensure_stage(size_t elems) {
    if (ctx->stage_elems >= elems) { used[0] = used[1] = false; return; }
    for (int b = 0; b < 2; b++) {
        if (ctx->stage[b]) cudaFree(ctx->stage[b]);
        JCHECK(cudaMalloc(&ctx->stage[b], elems * sizeof(double)));
    }
    ctx->stage_elems = elems;
    used[0] = used[1] = false;
}

// owner: panel p of Mo is final -> pack rows >= c0 (and the panel's Dinv blocks) into slot p & 1
void StreamPanelComm::publish_panel(int p) {
    const int b = p & 1;
    const int64_t c0 = (int64_t)p * pw * kBlk;
    const int64_t cols = std::min<int64_t>((int64_t)pw * kBlk, np - c0), rows = np - c0;
    const int64_t lcol = (int64_t)(p / ctx->world) * pw * kBlk;        // first local column of own panel p
    if (used[b]) JCHECK(cudaStreamWaitEvent(compute, ctx->ev_bcast[b], 0));   // the previous transfer through this slot is done
    launch_copy2d(ctx->stage[b], cols, Mo + c0 * ldo + lcol, ldo, rows, cols, compute);
    launch_copy2d(ctx->stage[b] + rows * cols, kBlk, Dinv + c0 * kBlk, kBlk, cols, kBlk, compute);
    JCHECK(cudaEventRecord(ctx->ev_ready[b], compute));
}

StreamPanelComm::Ref StreamPanelComm::get_panel_impl(int k, int64_t c0, int64_t cols, int root) {
    const int b = k & 1;
    const int64_t rows = np - c0;
    const size_t count = (size_t)rows * cols + (size_t)cols * kBlk;
    if (ctx->rank == root) {
        JCHECK(cudaStreamWaitEvent(ctx->net, ctx->ev_ready[b], 0));
        nccl_check(g_nccl.Broadcast(ctx->stage[b], ctx->stage[b], count, kNcclFloat64, root, ctx->comm, ctx->net), "ncclBroadcast");
        JCHECK(cudaEventRecord(ctx->ev_bcast[b], ctx->net));
    } else {
        if (used[b]) JCHECK(cudaStreamWaitEvent(ctx->net, ctx->ev_unpacked[b], 0));   // the slot's previous panel has been consumed
        nccl_check(g_nccl.Broadcast(ctx->stage[b], ctx->stage[b], count, kNcclFloat64, root, ctx->comm, ctx->net), "ncclBroadcast");
        JCHECK(cudaEventRecord(ctx->ev_bcast[b], ctx->net));
        JCHECK(cudaStreamWaitEvent(compute, ctx->ev_bcast[b], 0));
        launch_copy2d(Dinv + c0 * kBlk, kBlk, ctx->stage[b] + rows * cols, kBlk, cols, kBlk, compute);
    }
    used[b] = true;
    root_of[b] = root;
    // virtual base: L(r, c) = base[r * ld + c] for global r >= c0, c in [c0, c0 + cols)
    return Ref{ctx->stage[b] - c0 * cols - c0, cols};
}

void StreamPanelComm::done_panel(int k) {
    const int b = k & 1;
    if (root_of[b] != ctx->rank) JCHECK(cudaEventRecord(ctx->ev_unpacked[b], compute));
}

void StreamPanelComm::phase_boundary() {
    if (ev_phase) JCHECK(cudaEventRecord(ev_phase, compute));
}

}  // namespace jaicov
