// internal.h — device-side data layout and kernel launchers shared by the .cu files.
//
// HBM layout (per cloud, working frame = normalised coordinates for the SE(3) entries, raw for run_icp):
//   x,y,z        FP64 SoA, ORIGINAL point order            (gathers by correspondence index)
//   sx,sy,sz     FP64 SoA, MORTON order                    (coalesced leaf loads in the traversals)
//   perm         int32  morton position -> original index
//   keys         uint64 sorted 63-bit Morton codes
//   box          float  [6][nodes] outward-rounded AABBs of the implicit 32-wide hierarchy
//                (level 0 = leaves of 32 consecutive Morton points, level l+1 = 32 nodes of level l)
//   frame        FP64 [9][n]  TOLDI rotation, 12-vector order (x-axis, y-axis, z-axis), original order
//   nrm / cov    FP64 [3][n] / [6][n] (symmetric packing 00,01,02,11,12,22), original order
//   rows32       float4 [3][n] 12-float SE(3) rows (alpha R | beta p), MORTON order   (brute-force tiles)
//   rows64       FP64 [12][n] same rows, MORTON order                                  (exact evaluation)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/se3icp.h"

namespace se3 {

constexpr int kMaxLevels = 8;
constexpr int kReducePartials = 32;  // doubles per block partial record
constexpr int kReduceBlocks = 296;   // 2 x 148 SMs: fixed grid -> deterministic partial layout

struct CloudIndex {
    int n = 0;
    int n_levels = 0;
    int total_nodes = 0;
    int start_level = 0;  // where traversals begin (see IndexStorage::plan_levels)
    int level_off[kMaxLevels] = {0};
    int level_cnt[kMaxLevels] = {0};
    const double* x = nullptr;
    const double* y = nullptr;
    const double* z = nullptr;
    const double* sx = nullptr;
    const double* sy = nullptr;
    const double* sz = nullptr;
    const int* perm = nullptr;      // Morton position -> original index
    const int* inv = nullptr;       // original index -> Morton position
    const uint64_t* keys = nullptr;
    const float2* box = nullptr;    // [3][total_nodes] (lo, hi) per axis, rounded outwards
    const double* bbox = nullptr;   // device: lo[3], hi[3] of the cloud (for Morton quantisation)
};

// Everything the iteration kernels need that changes on the device between iterations.
struct IterState {
    double T_total[16];
    double T_prev[16];
    double T_i[16];
    double T_final[16];
    double c_src[3];
    double c_tgt[3];
    double scale;
    double mse_prev, mse_cur, mse_rel, T_change;
    double tgt_absmax;  // max |coordinate| over the target SE(3) rows (certification bound)
    int iter;
    int se3_iters;
    int switch_icp;
    int done;
    int n_keep;
    unsigned int thr_bits;  // trimmed rejection: key of the n_keep-th correspondence (float bits, complemented when keeping the largest)
    int eq_budget;          // ... how many of the correspondences whose key equals thr_bits survive
    int tie_limit;          // ... which: those with source index <= tie_limit (index order, like the mask kernels)
    int peer_timeout;       // sharded pair: a peer's record did not arrive within kPeerTimeoutNs (run aborted)
    int repair_count;
    int hist_count;
    int switch_iter;  // value of iter when the ICP phase began (-1 before)
    int work_count;   // entries of the coherence work list (reset every iteration)
    int corr_stamped; // the end of this iteration's correspondence stage has been time-stamped
    unsigned long long searched_total;  // queries the coherence filter could NOT settle, summed over the iterations
    long long total_repairs;
    unsigned long long t_mark;      // globaltimer at the end of the previous solve/update
    unsigned long long t_corr_ns;   // accumulated correspondence-search time (both phases)
    unsigned long long t_corr_se3_ns;  // ... of which SE(3)-phase iterations (filter + 12-D search kernels)
    unsigned long long t_start;     // globaltimer at state initialisation
    unsigned long long t_switch;    // globaltimer when the ICP phase began
};

// Fixed (per run) configuration passed by value to the kernels.
struct RunConfig {
    int entry;
    int variant;
    int max_iter;
    int max_se3_iter;
    int has_se3;      // entry != RUN_ICP
    int pure;         // entry == RUN_SE3_PURE
    int with_cf;      // entry == RUN_SE3_ICP_CF
    int trim_active;  // floor(float(overlap) * N) < N
    int n_keep_target;
    int keep_largest;
    int record_history;
    int max_history;
    int coherence;       // SE(3) search may skip queries whose remembered match is provably still the nearest
    int coherence_xyz;   // same for the 3-D search of the ICP phase / run_icp
    double coherence_thr;  // ... once ||T_prev - T_total||_F of the last iteration is below this
    double reseed_thr;     // unsettled queries also try a fresh Morton seed while that change is above this
    double mse;
    double mse_switch;
    double alpha;
    double beta;
};

struct SourceView {
    int n;                // points in the cloud (plane stride of frame / cov)
    int begin, end;       // query range owned by this rank: [0, n) unless the pair is sharded
    const int* order;     // optional processing order (a permutation of [0, n): the source's Morton order), or null
    const double* x;      // working-frame original coordinates p0 (never rewritten: q = T_total * p0 on the fly)
    const double* y;
    const double* z;
    const double* frame;  // [9][n]
    const double* cov;    // [6][n] or null
    const double* conf;   // [n] or null
};

struct TargetView {
    int n;
    CloudIndex idx;
    const double* nrm;    // [3][n] or null
    const double* cov;    // [6][n] or null
    const double* conf;   // [n] or null
    // what the reduction gathers per correspondence, as ONE 96-byte record per target point (3 sectors instead of the
    // 3 + 3 / 3 + 6 sectors of the plane layout): x y z | normal(3) or covariance 00 01 02 11 12 22 | padding; or null
    const double* rec;
    // SE(3) search structure: rows in 6-D Morton order (se3_index.cu); level layout shared with idx
    const float4* rows32; // [3][n]  (alpha R | tscale p) as floats
    const double* rows64; // [12][n]
    const float2* box12;  // (lo, hi) per dimension, rounded outwards; two dimensions per 16-byte word: box12_slot()
    const int* perm12;    // 6-D position -> original index
    const int* inv12;     // original index -> 6-D position
    const uint64_t* keys12;
    double tscale;        // scale of the translation part of the rows (beta; 1 for run_se3_icp_with_cf, .cpp:834-836)
    double dist_scale;    // beta / tscale: stored distance uses beta * p (.cpp:465)
};

constexpr int kTcountWords = 4 + 64;
constexpr int kTrimHistBins = 65536;  // histogram of the top 16 bits of the distance keys of one iteration

// 12-D boxes: dimensions 2 j and 2 j + 1 of a node share one float4 (lo, hi, lo, hi) in plane j of 6, so that a node test
// is six 16-byte loads.  Index of the float2 holding dimension k of `node`:
__host__ __device__ inline size_t box12_slot(int k, int node, size_t total_nodes) {
    return ((size_t)(k >> 1) * total_nodes + (size_t)node) * 2 + (size_t)(k & 1);
}

struct CorrBuffers {
    double* d2_nd;   // optional [N] squared distance in the search space (12-D or 3-D), stage API
    int* idx;        // [N] matched target (original index), persists across iterations (warm start)
    double* dist;    // [N] FP64 distance (reference distances_vec)
    float* distf;    // [N] stored float distance (pcl::Correspondence::distance)
    uint8_t* keep;   // [N] trim mask (valid when trim_active)
    int* repair;     // [N] queries needing the exact FP64 repair
    int* work;       // [N] queries the coherence filter could not settle this iteration
    // coherence filter: the query a remembered second-nearest distance belongs to is T_ref * X0, and T_ref is the estimate
    // of the iteration that recorded it — so the iteration number is remembered (4 bytes) instead of the 12-vector
    // (96 bytes written per searched query, read per filtered query), and the estimates of all iterations are kept
    int* ref_iter;         // [N]
    double* t_table;       // [t_table_cap][16] estimate the queries of iteration k were formed with
    int t_table_cap;
    double* ref_d2nd;  // [N] exact distance to the second-nearest row at that time, < 0 = not known
    // single-pass trimmed rejection (single-GPU runs): the kernels that store a distance also count its key in
    // thist (non-null only then); trim_select turns the histograms into (thr_bits, tie_limit)
    unsigned int* thist;        // [kTrimHistBins]
    unsigned long long* tcand;  // [N] (key << 32 | source index) of the correspondences in the threshold's 16-bit bin
    unsigned int* tcount;       // [kTcountWords]: candidate count, block tickets, coarse histogram (optimise.cu)
};

// ---- launchers (all asynchronous on `st`) ------------------------------------------------------
// spatial_index.cu
struct IndexStorage;  // owns the buffers behind a CloudIndex
int launch_sum_xyz(const double* aos, int n, double* partial /*[kReduceBlocks*3]*/, cudaStream_t st);
int launch_maxdist(const double* aos, int n, const double* partial_sum, double* partial_max /*[kReduceBlocks]*/,
                   cudaStream_t st);
int launch_normalise(const double* aos, int n, const double* psum_self, const double* pmax_src,
                     const double* pmax_tgt, int n_src, int n_tgt, double scale_pre, int which, IterState* state,
                     double* x, double* y, double* z, cudaStream_t st);
int launch_aos_to_soa(const double* aos, int n, double* x, double* y, double* z, cudaStream_t st);
int launch_confidence(const double* aos, int n, double* conf, cudaStream_t st);

// knn_features.cu
struct FeatureArgs {
    int k_lrf;        // 0 = no LRF
    int k_nrm;        // 0 = no normals
    int want_cov;     // GICP covariance from the normal
    double gicp_eps;
    double* frame;    // [9][n]
    double* nrm;      // [3][n]
    double* cov;      // [6][n]
    int* knn_idx;     // optional [n*K] (stage API), original order rows
    double* knn_d2;   // optional [n*K]
    int K;            // list length = max(k_lrf, k_nrm, requested)
    int q_begin, q_end;  // only points whose ORIGINAL index lies in [q_begin, q_end) are processed (sharded source)
    // partial range: the Morton positions of those points, compacted by launch_knn_features itself (scratch of n ints
    // + 1 counter), so that the warps of a block all have work wherever the range lies in the Morton order
    int* active_list;
    int* active_count;
};
int launch_knn_features(const CloudIndex& I, const FeatureArgs& fa, cudaStream_t st);
// shot_lrf.cu: frame planes [9][n] as the feature pass writes them; *unresolved counts median votes that could not be decided
int launch_shot_lrf(const CloudIndex& I, double radius, double* frame, int* unresolved, cudaStream_t st);
int launch_cov_from_normals(const double* nrm /*[3][n]*/, int n, double eps, double* cov /*[6][n]*/, cudaStream_t st);

// nn_search.cu
int launch_nn_filter(const SourceView& S, const TargetView& T, const RunConfig& cfg, IterState* state, CorrBuffers cb,
                     cudaStream_t st);
int launch_nn_search(const SourceView& S, const TargetView& T, const RunConfig& cfg, IterState* state, CorrBuffers cb,
                     cudaStream_t st);  // 12-D or 3-D search, chosen on the device by the phase flag
int launch_nn_se3_tree(const SourceView& S, const TargetView& T, const RunConfig& cfg, IterState* state, CorrBuffers cb,
                       cudaStream_t st);
int launch_nn_se3_brute(const SourceView& S, const TargetView& T, const RunConfig& cfg, IterState* state,
                        CorrBuffers cb, int force_all_repair, cudaStream_t st);
int launch_nn_se3_repair(const SourceView& S, const TargetView& T, const RunConfig& cfg, IterState* state,
                         CorrBuffers cb, cudaStream_t st);
int launch_nn_xyz(const SourceView& S, const TargetView& T, const RunConfig& cfg, IterState* state, CorrBuffers cb,
                  cudaStream_t st);

// optimise.cu
int launch_trim(const RunConfig& cfg, IterState* state, CorrBuffers cb, int n, unsigned int* hist /*[4*256]*/,
                int* block_eq /*[kReduceBlocks]*/, cudaStream_t st);
int launch_trim_hist(const RunConfig& cfg, IterState* state, const float* distf, int n, unsigned int* hist, int pass,
                     cudaStream_t st);
int launch_trim_count_eq(const RunConfig& cfg, IterState* state, const float* distf, int n, const unsigned int* hist,
                         int* block_eq, int* eq_total, cudaStream_t st);
int launch_trim_apply(const RunConfig& cfg, IterState* state, const float* distf, int n, const unsigned int* hist,
                      const int* block_eq, const int* rank_eq, int rank, uint8_t* keep, cudaStream_t st);
// One very large pair sharded over the GPUs of a node: the per-iteration all-reduce of the 29-double normal-equation
// record runs INSIDE the iteration's last kernel over peer memory (NVLink / NVSwitch).  Every rank owns a mailbox of
// kPeerSlotWords-word slots [2 parities][world]; the last block of reduce_kernel stores its rank's record into slot
// [parity][rank] of every peer's mailbox, publishes it with a sequence word, waits until the world's records of this
// iteration have arrived in its own mailbox and sums them in rank order — identical bits on every rank, no host
// round trip, no library call, and the loop stays one CUDA graph.
constexpr int kPeerSlotWords = 32;  // 31 doubles of record + 1 sequence word (256 bytes)
constexpr int kMaxPeers = 16;
struct PeerReduce {
    unsigned long long* const* mailboxes;  // device array [world]: every rank's mailbox, peer-mapped (own entry: local)
    int world, rank;                       // world <= 1: no exchange
    unsigned long long seq_base;           // run counter << 32 (identical on all ranks); sequence = seq_base + iteration
};

// the tail of an iteration folded into reduce_kernel (its last block to finish runs the solve / update / stop logic)
struct SolveFusion {
    int enabled;
    PeerReduce peer;
    double* history;                 // per-iteration T_i or null
    unsigned int* hist;              // histograms of the multi-pass trim to clear for the next iteration, or null
    unsigned long long cond_handle;  // cudaGraphConditionalHandle of the loop graph, 0 = none
};
int launch_reduce(const SourceView& S, const TargetView& T, const RunConfig& cfg, IterState* state, CorrBuffers cb,
                  double* partials /*[kReduceBlocks*kReducePartials]*/, const SolveFusion& fuse, cudaStream_t st);
int launch_trim_select(const RunConfig& cfg, IterState* state, CorrBuffers cb, int begin, int end, cudaStream_t st);
int launch_trim_stage(const RunConfig& cfg, IterState* state, CorrBuffers cb, int n, cudaStream_t st);
constexpr int kTargetRecordDoubles = 12;
int launch_pack_target_records(const CloudIndex& I, const double* nrm, const double* cov, double* rec, cudaStream_t st);
int launch_sum_partials(const double* partials /*[kReduceBlocks][kReducePartials]*/, double* total /*[kReducePartials]*/,
                        cudaStream_t st);
int launch_solve_update(const RunConfig& cfg, IterState* state, const double* partials, int n_records, double* history,
                        unsigned int* hist /*[4*256] trim histograms, zeroed for the next iteration*/,
                        unsigned long long cond_handle /*0 = none*/, cudaStream_t st);
int launch_loop_condition(unsigned long long cond_handle, const IterState* state, cudaStream_t st);
int launch_finalize(const RunConfig& cfg, IterState* state, cudaStream_t st);
int launch_init_state(IterState* state, unsigned int* hist, cudaStream_t st);
int launch_mark_loop_start(IterState* state, cudaStream_t st);

}  // namespace se3
