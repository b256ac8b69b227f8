/*
 * se3icp.h — C ABI of the B200-native SE(3)-ICP registration path (libse3icp_cuda.so).
 *
 * The reference (kenahm/se3-icp) has no FFI: its boundary is the C++ class
 * IterativeSE3Registration (reference include/iterative_SE3_registration.hpp:27-99) whose
 * run_*() methods do all the work on the CPU.  This header is the thin C layer the re-implemented
 * host class (include/iterative_SE3_registration.hpp in this repo) calls instead; every entry
 * point names the reference function it replaces.  Plain pointers and sizes only; every function
 * returns an int status (SE3ICP_OK == 0) and never throws.  Host buffers in, host buffers out
 * unless a name says "_device".  A context owns all device memory and one CUDA stream; a context
 * is not thread-safe, different contexts are independent.
 *
 * There is no CPU fallback: se3icp_create fails with SE3ICP_ERR_NO_DEVICE when no sm_100 GPU is
 * usable.
 */
#ifndef SE3ICP_H
#define SE3ICP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SE3ICP_ABI_VERSION 4

enum se3icp_status {
    SE3ICP_OK = 0,
    SE3ICP_ERR_ARG = 1,        /* null pointer / bad size / bad enum */
    SE3ICP_ERR_NO_DEVICE = 2,  /* no usable CUDA device */
    SE3ICP_ERR_CUDA = 3,       /* a CUDA call failed; see se3icp_last_error */
    SE3ICP_ERR_NCCL = 4,
    SE3ICP_ERR_UNSUPPORTED = 5, /* e.g. kNN k larger than SE3ICP_MAX_KNN */
    SE3ICP_ERR_STATE = 6        /* clouds not set, etc. */
};

#define SE3ICP_MAX_KNN 128

/* optimisation step, reference strings "pt2pt" | "pt2pl" | "gicp"
 * (reference src/iterative_SE3_registration.cpp:524-531, 691-698) */
enum se3icp_variant { SE3ICP_PT2PT = 0, SE3ICP_PT2PL = 1, SE3ICP_GICP = 2 };

/* which run_* method of the reference class is replaced */
enum se3icp_entry {
    SE3ICP_RUN_ICP = 0,        /* run_icp            .cpp:473-552  */
    SE3ICP_RUN_SE3_ICP = 1,    /* run_se3_icp        .cpp:555-739  */
    SE3ICP_RUN_SE3_ICP_CF = 2, /* run_se3_icp_with_cf .cpp:742-959 */
    SE3ICP_RUN_SE3_PURE = 3    /* run_se3_pure       .cpp:962-1128 */
};

/* SE(3) nearest-neighbour strategy */
enum se3icp_nn_mode {
    SE3ICP_NN_AUTO = 0,       /* library picks per iteration */
    SE3ICP_NN_BRUTE_F32 = 1,  /* tiled FP32 brute force + certification + exact FP64 repair */
    SE3ICP_NN_EXACT_F64 = 2,  /* FP64 brute force for every query (slow; test oracle on device) */
    SE3ICP_NN_TREE = 3        /* pruned traversal of the 12-D bounded hierarchy, FP64 leaves */
};

enum se3icp_which { SE3ICP_SOURCE = 0, SE3ICP_TARGET = 1 };

/* local reference frame that lifts a point to an SE(3) element */
enum se3icp_lrf_method {
    SE3ICP_LRF_TOLDI = 0, /* kNN support, .cpp:241-331: what every entry point of the reference runs */
    SE3ICP_LRF_SHOT = 1   /* radius support, .cpp:121-239: the reference's dormant alternative (its calls are commented out
                             at .cpp:593-594,812-813); LRF ablations */
};

/* POD mirror of the public configuration fields of IterativeSE3Registration
 * (reference hpp:80-95; defaults .cpp:334-348). */
typedef struct se3icp_params {
    int32_t variant;                /* se3icp_variant */
    int32_t entry;                  /* se3icp_entry */
    int32_t max_num_iterations;     /* 150 */
    int32_t max_num_se3_iterations; /* 20 */
    int32_t number_of_nn_for_LRF;   /* 30 */
    int32_t knn_normals_pt2pl;      /* 30 (Open3D EstimateNormals default, .cpp:494,643) */
    int32_t knn_normals_gicp;       /* 20 (.cpp:43) */
    int32_t trim_keep_largest;      /* 1 (default) = what PCL 1.14 does: CorrespondenceRejectorTrimmed runs std::nth_element
                                       with pcl::isBetterCorrespondence, which is `pc1.distance > pc2.distance`, so the
                                       floor(overlap * N) correspondences with the LARGEST distances survive.  0 = keep the
                                       smallest (what the class documentation says it does).  PCL is not installed here, so
                                       the comparator is restated from its published source, not verified */
    double mse;                     /* 1e-5 */
    double mse_switch_error;        /* 1e-3 */
    double estimated_overlap;       /* 1.0 */
    double alpha_rot;               /* 3.0 */
    double beta_transl;             /* 1.0 */
    double scale_preprocessing;     /* 3.0 */
    double gicp_epsilon;            /* 1e-3 (.cpp:498-499,646-647) */
    int32_t nn_mode;                /* se3icp_nn_mode */
    int32_t use_graph;              /* 1 = capture the iteration in a CUDA graph */
    int32_t record_history;         /* 1 = keep per-iteration T_i (reference estimated_history_, hpp:63) */
    int32_t nn_coherence;           /* 1 = SE(3) search may skip queries whose previous match is provably still nearest */
    int32_t reuse_features;         /* 1 = keep a cloud's LRFs / normals / covariances across runs while the cloud stays in the
                                       context (se3icp_swap_clouds, se3icp_run_sequence); 0 = recompute every run.
                                       se3icp_set_cloud / se3icp_set_cloud_device always invalidate them, so a caller who
                                       rewrites a caller-owned device buffer in place must set the cloud again */
    int32_t lrf_method;             /* se3icp_lrf_method; 0 = TOLDI (default) */
    int32_t reserved0;
    double lrf_radius;              /* 0.8 (reference lrf_radius_, .cpp:340): support radius of the SHOT frame, in the
                                       NORMALISED cloud's units (.cpp:568-582), as the commented call sites pass it.  SHOT
                                       frames depend on that per-pair scale, so they are never reused across runs, and
                                       se3icp_run_sharded does not take them */
} se3icp_params;

typedef struct se3icp_stats {
    int32_t num_iterations;          /* reference num_iterations_ */
    int32_t num_pure_se3_iterations; /* reference num_pure_se3_iterations_ (-1 for run_icp) */
    double scaling_factor;           /* s = scale_preprocessing / r_max (.cpp:574) */
    double time_total_ms;            /* device time of the whole run (CUDA events) */
    double time_setup_ms;            /* normalise + index + LRF + normals/covariances */
    double time_se3_correspondence_search_ms; /* reference time_se3_correspondence_search_ (hpp:86) */
    double time_before_pure_icp_ms;  /* reference time_before_pure_icp_ (hpp:85) */
    int64_t exact_repairs;           /* queries re-done by the exact FP64 repair kernel */
    int64_t kernel_launches;         /* kernels of this library launched during the run */
    double time_se3_phase_search_ms; /* device time of the 12-D correspondence stage, summed over the SE(3) iterations */
    int64_t feature_reuses;          /* clouds whose neighbourhood features were taken from an earlier run (0, 1 or 2) */
    int64_t queries_searched;        /* nearest-neighbour queries that ran a tree search, summed over the iterations; the
                                        rest (iterations x source points - this) were settled by the coherence filter */
    int64_t graph_instantiations;    /* loop-graph executables this context has created so far (it keeps one and
                                        re-parameterises it from run to run; a count that grows with the runs means the
                                        launch sequence keeps changing) */
    int64_t loop_was_graph;          /* 1 = this run's iteration loop was one CUDA graph launch; 0 = host-driven (one
                                        synchronisation per iteration: use_graph = 0, a process under Nsight Compute, or a
                                        sharded pair that all-reduces through NCCL) */
} se3icp_stats;

typedef struct se3icp_ctx se3icp_ctx;

int se3icp_abi_version(void);
const char* se3icp_last_error(void);
void se3icp_default_params(se3icp_params* p);

/* device = CUDA ordinal.  stream = a cudaStream_t created by the caller (e.g. torch's current
 * stream) or NULL to let the context create its own. */
int se3icp_create(int device, void* stream, se3icp_ctx** out);
int se3icp_destroy(se3icp_ctx* ctx);
int se3icp_synchronize(se3icp_ctx* ctx);

/* replaces setSourceCloud / setTargetCloud (.cpp:350-376): copies xyz only; append != 0 mimics the
 * push_back behaviour of the PointCloud overloads. */
int se3icp_set_cloud(se3icp_ctx* ctx, int which, const double* xyz_aos, size_t n, int append);
/* same, from a device buffer already resident in HBM (AoS doubles) */
int se3icp_set_cloud_device(se3icp_ctx* ctx, int which, const double* d_xyz_aos, size_t n);

/* replaces run_icp / run_se3_icp / run_se3_icp_with_cf / run_se3_pure.  T_out row-major 4x4. */
int se3icp_run(se3icp_ctx* ctx, const se3icp_params* p, double* T_out, se3icp_stats* stats);
/* asynchronous form: enqueue everything on the context's stream, read back later.  Truly asynchronous only with
 * params.use_graph (the default): the host-driven loop (use_graph = 0, or a process running under Nsight Compute)
 * synchronises once per iteration inside se3icp_run_async.  While a run is pending every entry point that would
 * reallocate or overwrite its buffers (set_cloud*, swap_clouds, run_async, run_sharded, the stage-level calls)
 * returns SE3ICP_ERR_STATE until se3icp_run_finish has been called. */
int se3icp_run_async(se3icp_ctx* ctx, const se3icp_params* p);
int se3icp_run_finish(se3icp_ctx* ctx, double* T_out, se3icp_stats* stats);

/* per-iteration estimates T_i (reference estimated_history_); returns count via *n_out */
int se3icp_get_history(se3icp_ctx* ctx, double* T_hist, int max_entries, int* n_out);
/* final correspondences (reference current_correspondences_set, hpp:74) */
int se3icp_get_correspondences(se3icp_ctx* ctx, int32_t* tgt_idx, double* dist, size_t n);
/* SE(3) clouds as 4x4 row-major matrices (reference source_se3_cloud_/target_se3_cloud_, hpp:59-60) */
int se3icp_get_se3_cloud(se3icp_ctx* ctx, int which, double* frames16, size_t n);

/* batch of independent pairs on one GPU (reference benchmark_kitti.cpp:120-197 driver loop):
 * pair p uses src[p] (n_src[p] points) and tgt[p]; n_ctx contexts are cycled.  Host buffers. */
int se3icp_run_batch(se3icp_ctx** ctxs, int n_ctx, int n_pairs, const double* const* src, const size_t* n_src,
                     const double* const* tgt, const size_t* n_tgt, const se3icp_params* p, double* T_out /*[n_pairs*16]*/,
                     se3icp_stats* stats /*[n_pairs] or NULL*/);
/* same with inputs already on the device */
int se3icp_run_batch_device(se3icp_ctx** ctxs, int n_ctx, int n_pairs, const double* const* d_src, const size_t* n_src,
                            const double* const* d_tgt, const size_t* n_tgt, const se3icp_params* p, double* T_out,
                            se3icp_stats* stats);

/* Odometry-style sequences (reference examples/benchmark_kitti.cpp:120-131 registers scan i+1 onto scan i, so every
 * scan is the source of one pair and the target of the next).  se3icp_swap_clouds exchanges the two slots of the
 * context together with everything derived from them; with params.reuse_features the next run then skips the
 * kNN-90 / LRF / normal / covariance stage for the cloud that already went through it.  Those features are invariant
 * under the per-pair normalisation (a uniform scale about the cloud's own centroid, .cpp:568-582), so the result
 * agrees with an independent run of the pair to rounding (tests: <= 1e-9, equal iteration counts); it is not
 * bit-identical, because the reference recomputes the neighbourhoods at the new scale.  Extension, not in the reference. */
int se3icp_swap_clouds(se3icp_ctx* ctx);
/* registers scans[i+1] (source) onto scans[i] (target) for i = 0 .. n_scans-2, reusing each scan's features once.
 * scans[i]: n_points[i] x 3 doubles, host buffers (device_inputs = 0) or device buffers that stay valid (1). */
int se3icp_run_sequence(se3icp_ctx* ctx, const double* const* scans, const size_t* n_points, int n_scans,
                        const se3icp_params* p, int device_inputs, double* T_out /*[(n_scans-1)*16]*/,
                        se3icp_stats* stats /*[n_scans-1] or NULL*/);

/* one very large pair (BASELINE.json configs[4]): every rank holds both clouds (se3icp_set_cloud), owns the
 * source query range [src_begin, src_end) for LRF set-up, correspondence search and reduction, and the
 * 29-double normal-equation record is all-reduced once per iteration.  All ranks return the identical transform.
 * The communicator is either created by the library (se3icp_comm_unique_id on one rank, broadcast the
 * SE3ICP_COMM_ID_BYTES by any means, se3icp_comm_init on every rank; pass nccl_comm = NULL) or supplied by
 * the caller as an ncclComm_t with its rank / size.  NCCL is resolved at run time (dlopen libnccl.so.2).
 * se3icp_comm_init (and the first se3icp_run_sharded with a caller-owned communicator) is collective: besides the
 * communicator it sets up one 256-byte-per-rank mailbox per context, exchanged as CUDA IPC handles, through which the
 * ranks of one node all-reduce the record over NVLink peer memory INSIDE the iteration's last kernel — no host round
 * trip, the loop stays one CUDA graph.  When peer access is not available, when SE3ICP_SHARDED_P2P=0 is set, or when
 * the trimmed rejection is active (its 4 x 256-bin histograms are summed across ranks between passes), the record goes
 * through ncclAllReduce from a host-driven loop instead.  A rank whose peers do not show up within 20 s returns
 * SE3ICP_ERR_NCCL.  se3icp_run_sharded calls must be made by all ranks, the same number of times.
 * Contiguous ranges of a scan (image bands, laser rings) cost the ranks different amounts of search work; permuting the
 * source before upload so that every range samples the whole scan (Python: sharding.dealt_order) balances them. */
#define SE3ICP_COMM_ID_BYTES 128
int se3icp_comm_unique_id(void* id_out);
int se3icp_comm_init(se3icp_ctx* ctx, int n_ranks, int rank, const void* id);
int se3icp_comm_destroy(se3icp_ctx* ctx);
/* rank / size of the communicator the context currently holds (0 / 1 without one) */
int se3icp_comm_info(se3icp_ctx* ctx, int* rank_out, int* n_ranks_out);
int se3icp_run_sharded(se3icp_ctx* ctx, const se3icp_params* p, size_t src_begin, size_t src_end, void* nccl_comm,
                       int rank, int n_ranks, double* T_out, se3icp_stats* stats);

/* measurement hook (bench.py roofline): re-launches one kernel of the hot path `repeats` times on the
 * context's stream, on the device data left by the last se3icp_run (at its final pose), and returns the average
 * launch duration measured with CUDA events on that stream.  The two searches are timed COLD: every query, no
 * remembered match, no coherence shortcut.  The context's correspondences are overwritten. */
enum se3icp_stage { SE3ICP_STAGE_NN_SE3 = 0, SE3ICP_STAGE_NN_XYZ = 1, SE3ICP_STAGE_REDUCE = 2, SE3ICP_STAGE_KNN_TARGET = 3 };
int se3icp_time_stage(se3icp_ctx* ctx, int stage, int repeats, double* ms_avg);

/* ---------------- stage-level entry points (parity tests; one per row of SURVEY §8a) ----------------
 * They borrow the context's cloud slots and work buffers: after any of them the context holds NO clouds (a later
 * se3icp_run without se3icp_set_cloud returns SE3ICP_ERR_STATE) and se3icp_get_* no longer refer to an earlier run. */

/* a3+a4: exact kNN of every point in its own cloud, ascending by (d2, index).  idx[n*k], d2[n*k] */
int se3icp_knn(se3icp_ctx* ctx, const double* xyz, size_t n, int k, int32_t* idx, double* d2);
/* a4: TOLDI LRF (.cpp:241-331).  frames[n*16] row-major [x y z p; 0 0 0 1] */
int se3icp_lrf(se3icp_ctx* ctx, const double* xyz, size_t n, int k, double* frames);
/* f4: SHOT LRF with radius support (.cpp:121-239, the reference's dormant alternative to TOLDI; `lrf_radius_` .cpp:340).
 * frames[n*16] as se3icp_lrf.  A point with fewer than 5 others inside the radius gets the identity rotation (undefined
 * in the reference).  *unresolved_ties (may be NULL): median votes (.cpp:189-197) that could not be decided because more
 * than ~100 support points share one distance; 0 on real data. */
int se3icp_shot_lrf(se3icp_ctx* ctx, const double* xyz, size_t n, double radius, double* frames, int64_t* unresolved_ties);
/* a6: Open3D EstimateNormals(KNN(k)).  normals[n*3], unoriented */
int se3icp_normals(se3icp_ctx* ctx, const double* xyz, size_t n, int k, double* normals);
/* a6: GICP covariances from normals (.cpp:4-14,45-51).  cov[n*9] row-major */
int se3icp_gicp_cov(se3icp_ctx* ctx, const double* normals, size_t n, double eps, double* cov);
/* a7: 12-D 1-NN (.cpp:444-470).  rows are [alpha R col-major(9), beta p(3)]; idx[n], d2[n] (12-D squared) */
int se3icp_nn_se3(se3icp_ctx* ctx, const double* src_rows, size_t n, const double* tgt_rows, size_t m, int nn_mode,
                  int32_t* idx, double* d2, int64_t* exact_repairs);
/* a8: 3-D 1-NN (.cpp:402-416) */
int se3icp_nn_xyz(se3icp_ctx* ctx, const double* queries, size_t n, const double* tgt_xyz, size_t m, int32_t* idx,
                  double* d2);
/* a9: trimmed rejection (PCL CorrespondenceRejectorTrimmed).  keep[n] = 0/1 */
int se3icp_trim(se3icp_ctx* ctx, const float* dist, size_t n, double overlap, int keep_largest, uint8_t* keep,
                int64_t* n_keep);
/* a11: Umeyama sums + closed-form solve on the given pairs -> T row-major */
int se3icp_reduce_pt2pt(se3icp_ctx* ctx, const double* src, size_t n, const double* tgt, size_t m,
                        const int32_t* corr_tgt /*[n], -1 = rejected*/, double* T_out);
/* a12: out27 = 21 upper-triangular JTJ entries (row-major) then 6 JTr */
int se3icp_reduce_pt2pl(se3icp_ctx* ctx, const double* src, size_t n, const double* tgt, const double* tgt_normals,
                        size_t m, const int32_t* corr_tgt, double* out27);
/* a13 (+ a20 weights, may be NULL): covariances row-major 3x3 */
int se3icp_reduce_gicp(se3icp_ctx* ctx, const double* src, const double* src_cov, size_t n, const double* tgt,
                       const double* tgt_cov, size_t m, const int32_t* corr_tgt, const double* conf_src,
                       const double* conf_tgt, double* out27);
/* a12/a13 tail: LDLT solve of JTJ x = -JTr and the Euler update (Open3D) -> T row-major */
int se3icp_solve(se3icp_ctx* ctx, const double* in27, double* T_out);

/* ---------------- evaluation helpers either side of the path, on the device ----------------
 * What the reference's drivers do around a registration with its cc library and Open3D.  Host buffers in and out;
 * like the stage-level entry points they leave the context without clouds. */

/* cc::error_filterreg (src/cc.cpp:4-20): mean over the source of || T_gt p - T_est p ||.  T row-major 4x4. */
int se3icp_eval_error_filterreg(se3icp_ctx* ctx, const double* src_xyz, size_t n, const double* T_gt, const double* T_est,
                                double* error_out);
/* cc::compute_corrs_with_gt (src/cc.cpp:116-143): index of the target point nearest to T_gt * src[i], exact, ties to the
 * smaller index.  (cc::compute_nearest_neighbor_correspondences, cc.cpp:220-236, is se3icp_nn_xyz.) */
int se3icp_eval_corrs_with_gt(se3icp_ctx* ctx, const double* src_xyz, size_t n, const double* tgt_xyz, size_t m,
                              const double* T_gt, int32_t* tgt_idx);
/* cc::evaluate_LRF_quality (src/cc.cpp:63-88): mean of angularErrorSO3_alt (cc.cpp:39-61, degrees) between the rotation of
 * T_gt * source_SE3[pairs[p][0]] and that of target_SE3[pairs[p][1]]; frames as 4x4 row-major matrices (se3icp_lrf /
 * se3icp_get_se3_cloud layout).  per_pair_error_deg (n_pairs doubles) may be NULL. */
int se3icp_eval_lrf_quality(se3icp_ctx* ctx, const double* src_frames16, size_t n, const double* tgt_frames16, size_t m,
                            const double* T_gt, const int32_t* pairs /*[n_pairs][2]*/, size_t n_pairs, double* mean_error_deg,
                            double* per_pair_error_deg);
/* Open3D PointCloud::RandomDownSample as examples/benchmark_synthetic.cpp:100,150 uses it: (size_t)(n * ratio) points,
 * uniformly without replacement, in shuffled order.  The shuffle is a counter-based hash of (seed, index) sorted on the
 * device, so for a given seed the SUBSET differs from Open3D's mt19937 stream; deterministic for a given seed.
 * xyz_out ((n * ratio) x 3) and index_out may each be NULL; *n_out receives the count. */
int se3icp_random_downsample(se3icp_ctx* ctx, const double* xyz, size_t n, double sampling_ratio, uint64_t seed, double* xyz_out,
                             int32_t* index_out, size_t* n_out);

#ifdef __cplusplus
}
#endif
#endif
