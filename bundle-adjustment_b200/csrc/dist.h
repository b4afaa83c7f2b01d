// dist.h -- multi-GPU context (NCCL communicator, network stream, panel staging) of one handle.
#pragma once
#include "common.h"

namespace jaicov {

struct NcclUniqueId { char internal[128]; };

void nccl_unique_id(NcclUniqueId *out);
void nccl_comm_init_all(void **comms, int n, const int *devices);

struct DistContext {
    int rank = 0, world = 1;
    void *comm = nullptr;            // ncclComm_t
    cudaStream_t net = nullptr;      // high-priority stream of the panel broadcasts
    double *stage[2] = {nullptr, nullptr};
    size_t stage_elems = 0;
    cudaEvent_t ev_ready[2] = {nullptr, nullptr}, ev_bcast[2] = {nullptr, nullptr}, ev_unpacked[2] = {nullptr, nullptr};
    cudaEvent_t ev_tmp = nullptr;
    void init(int rank, int world, const NcclUniqueId &id);
    void adopt(int rank, int world, void *comm);
    void destroy();
    void allreduce_sum(double *buf, size_t count, cudaStream_t s);
};

// Comm concept of DenseSchedule::potrf_distributed
struct PanelComm {
    DistContext *ctx;
    cudaStream_t compute;
    double *M;
    int64_t ld, np;
    double *Dinv;
    int pw;                          // panel width in 128-tiles
    bool used[2] = {false, false};
    void ensure_stage(size_t elems);
    void panel_ready(int p);
    void bcast_panel(int k, int64_t row0, int64_t rows, int64_t col0, int64_t cols, int root);
};

// Comm concept of DenseSchedule::factor_solve_invert_streamed (owner-only storage): a published panel is packed into a staging
// slot and broadcast; the receivers USE it where it lands (no copy into a replica) and give the slot back with done_panel.
struct StreamPanelComm {
    DistContext *ctx;
    cudaStream_t compute;
    double *Mo;                      // own panels, full height, own tiles side by side
    int64_t ldo, np;
    double *Dinv;                    // np x 128, replicated (64 MB at config 5): inverses of the diagonal blocks
    int pw;                          // panel width in 128-tiles
    cudaEvent_t ev_phase = nullptr;  // recorded on the compute stream between the forward and the backward phase (stage timing)
    bool used[2] = {false, false};
    struct Ref { double *base; int64_t ld; };
    void ensure_stage(size_t elems);
    void publish_panel(int p);
    template <class PanelRef>
    PanelRef get_panel(int k, int64_t c0, int64_t cols, int root, PanelRef own) {
        Ref r = get_panel_impl(k, c0, cols, root);
        if (ctx->rank == root) return own;
        return PanelRef{r.base, r.ld};
    }
    Ref get_panel_impl(int k, int64_t c0, int64_t cols, int root);
    void done_panel(int k);
    void phase_boundary();
    int root_of[2] = {-1, -1};
};

}  // namespace jaicov
