// context.h — the object behind se3icp_ctx: one stream, all device memory of one registration pair.
#pragma once

#include <cuda_runtime.h>

#include "index_storage.h"
#include "internal.h"

struct se3icp_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;

    // clouds as handed over by the caller (AoS doubles on the device)
    se3::DeviceBuf raw[2];
    const double* raw_view[2] = {nullptr, nullptr};  // raw[w].ptr, or a caller-owned device buffer
    size_t n[2] = {0, 0};

    se3::IndexStorage index[2];
    se3::DeviceBuf frame[2], nrm[2], cov[2], conf[2];
    se3::Se3IndexStorage se3idx;
    se3::DeviceBuf corr_idx, corr_dist, corr_distf, keep, repair, d2_nd, ref_iter, t_table, ref_d2nd, work;
    int t_table_cap = 0;
    se3::DeviceBuf psum[2], pmax[2];
    se3::DeviceBuf partials, hist, block_eq, history;
    se3::DeviceBuf state;
    se3::DeviceBuf scratch;  // stage-level API staging
    se3::DeviceBuf totals, eq_total, rank_eq;  // sharded pair: all-reduced record, threshold-tie counts
    se3::DeviceBuf thist, tcand, tcount;       // single-launch trimmed rejection: key histogram, candidates, counters / tickets

    // one very large pair sharded over ranks (se3icp_run_sharded)
    void* comm = nullptr;  // ncclComm_t
    bool comm_owned = false;
    int comm_rank = 0, comm_size = 1;
    bool sharded = false;
    int shard_begin = 0, shard_end = 0;
    // peer-memory all-reduce of the sharded pair (internal.h: PeerReduce): own mailbox, the peers' mailboxes opened
    // through CUDA IPC, the device array of all of them, and the run counter that makes sequence words unique
    se3::DeviceBuf mailbox, mailbox_table;
    se3::DeviceBuf knn_list, knn_count;  // compacted query list of a partial kNN / feature pass
    se3::DeviceBuf eval_buf;             // staging of the evaluation entry points (eval.cu)
    se3::DeviceBuf tgt_rec;              // per-target-point gather record of the reduction (TargetView::rec)
    bool tgt_rec_valid = false;          // ... packed by the current run's set-up (stage-level calls use the planes)
    void* peer_ptr[se3::kMaxPeers] = {nullptr};
    bool peer_ready = false;
    unsigned long long peer_runs = 0;

    // What frame[w] / nrm[w] / cov[w] currently hold: the neighbourhood features of the cloud in slot w are invariant
    // under the per-pair normalisation (uniform scale about the cloud's own centroid), so a scan that was the source of
    // one pair can serve as the target of the next without recomputing them (se3icp_swap_clouds).
    struct FeatureKey {
        bool valid = false;
        size_t n = 0;
        int k_lrf = 0, k_nrm = 0, want_cov = 0;
        double eps = 0.0;
    } feat[2];
    long long feature_reuses = 0;

    se3::IterState* h_state = nullptr;  // pinned
    int* h_flag = nullptr;              // pinned
    cudaEvent_t ev_begin = nullptr, ev_setup = nullptr, ev_end = nullptr;

    se3::RunConfig cfg{};
    se3icp_params params{};
    bool run_pending = false;
    bool src_index_built = false;  // index[0].perm holds the source's Morton order of the current run
    cudaGraph_t loop_graph = nullptr;        // WHILE-loop graph of the last run (use_graph)
    cudaGraphExec_t loop_exec = nullptr;
    long long launches_per_iter = 0;
    long long graph_instantiations = 0, graph_updates = 0;  // loop graph: executables created / re-parameterised in place
    bool graph_run = false;
    bool variant_valid = true;
    long long launches = 0;
    int history_capacity = 0;

    se3::IterState* dstate() const { return state.as<se3::IterState>(); }
    se3::SourceView source_view() const;
    se3::TargetView target_view() const;
    se3::CorrBuffers corr_buffers(bool with_d2) const;
};
