// optimise.cu — trimmed rejection, normal-equation reduction, on-device solve and loop control
// (SURVEY §8 a9-a18).
//
//  * trim_select     PCL CorrespondenceRejectorTrimmed (reference .cpp:634-635,669-671) in one launch: the kernels
//                    that store a correspondence distance also count its key in a 16-bit histogram; one block turns
//                    it into the (key, source index) pair of the n_keep-th correspondence and the reduction applies
//                    `(key, index) <= threshold` on the fly.  Ties at the threshold are kept in index order.
//  * trim_hist/count_eq/apply  the same selection as four all-reducible 8-bit passes + a keep mask: the sharded pair,
//                    where the histogram has to be summed over the ranks between the passes.
//  * reduce          one fused gather + residual/Jacobian + FP64 accumulation pass per iteration for
//                    pt2pt (Umeyama sums, reference .cpp:692), pt2pl (.cpp:695) and GICP (.cpp:57-110,
//                    698, with the confidence weights of .cpp:913 for run_se3_icp_with_cf).  The source
//                    point and its covariance are T_total * p0 and R C0 R^T formed on the fly, so the
//                    reference's Transform() pass (.cpp:706) never touches memory.
//  * solve_update    fixed-order sum of the per-block partials, 6x6 LDL^T solve or 3x3 Kabsch, Euler
//                    update, T accumulation, mean-distance bookkeeping and the phase / stop logic of
//                    .cpp:709-729 (run_icp: .cpp:544-550, run_se3_pure: .cpp:1118), all on the device.
#include "common.cuh"
#include "distance_store.cuh"
#include "internal.h"

namespace se3 {

__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ bool se3_phase_active_o(const RunConfig& cfg, const IterState* st) {
    return cfg.has_se3 && (cfg.pure || !st->switch_icp);
}

// called by the first thread of the first kernel that follows the correspondence stage of an iteration
__device__ __forceinline__ void stamp_correspondence_end(const RunConfig& cfg, IterState* state) {
    if (state->corr_stamped) return;
    state->corr_stamped = 1;
    unsigned long long now = global_timer_ns();
    state->t_corr_ns += now - state->t_mark;
    if (se3_phase_active_o(cfg, state)) state->t_corr_se3_ns += now - state->t_mark;
}

// ------------------------------------------------------------------------------------------------
// state
// ------------------------------------------------------------------------------------------------
__global__ void init_state_kernel(IterState* st, unsigned int* hist) {
    int t = threadIdx.x;
    if (t < 16) {
        double v = (t % 5 == 0) ? 1.0 : 0.0;
        st->T_total[t] = v;
        st->T_prev[t] = v;
        st->T_i[t] = v;
        st->T_final[t] = v;
    }
    if (t == 0) {
        for (int k = 0; k < 3; k++) st->c_src[k] = st->c_tgt[k] = 0.0;
        st->scale = 1.0;
        st->mse_prev = st->mse_cur = st->mse_rel = 10000000.0;  // reference .cpp:485,631
        st->T_change = 10000000.0;
        st->tgt_absmax = 0.0;
        st->iter = 0;
        st->se3_iters = 0;
        st->switch_icp = 0;
        st->done = 0;
        st->n_keep = 0;
        st->thr_bits = 0;
        st->eq_budget = 0;
        st->tie_limit = 0x7fffffff;
        st->peer_timeout = 0;
        st->repair_count = 0;
        st->hist_count = 0;
        st->switch_iter = -1;
        st->work_count = 0;
        st->corr_stamped = 0;
        st->total_repairs = 0;
        st->searched_total = 0;
        st->t_corr_ns = 0;
        st->t_corr_se3_ns = 0;
        st->t_start = global_timer_ns();
        st->t_mark = st->t_start;
        st->t_switch = 0;
    }
    if (hist)
        for (int k = t; k < 4 * 256; k += blockDim.x) hist[k] = 0;
}

// ------------------------------------------------------------------------------------------------
// trimmed rejection
// ------------------------------------------------------------------------------------------------
// trim_key(): distance_store.cuh

// warp-cooperative: bin holding rank k in a 256-bin histogram; k becomes the rank inside that bin
__device__ __forceinline__ int warp_select_bin(const unsigned int* __restrict__ hist, unsigned int& k, int lane) {
    unsigned int loc[8], sum = 0;
#pragma unroll
    for (int e = 0; e < 8; e++) {
        loc[e] = hist[lane * 8 + e];
        sum += loc[e];
    }
    unsigned int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned int v = __shfl_up_sync(SE3_FULL, incl, o);
        if (lane >= o) incl += v;
    }
    unsigned int excl = incl - sum;
    bool mine = (k >= excl) && (k < incl);
    unsigned int m = __ballot_sync(SE3_FULL, mine);
    int owner = m ? (__ffs(m) - 1) : 31;
    int bin = 0;
    unsigned int krem = 0;
    if (lane == owner) {
        unsigned int c = excl;
        bin = lane * 8 + 7;
        krem = 0;
        for (int e = 0; e < 8; e++) {
            if (k < c + loc[e]) {
                bin = lane * 8 + e;
                krem = k - c;
                break;
            }
            c += loc[e];
        }
    }
    bin = __shfl_sync(SE3_FULL, bin, owner);
    k = __shfl_sync(SE3_FULL, krem, owner);
    return bin;
}

// chain of selections for passes [0, upto): returns the key prefix (high digits) and the remaining rank
__device__ __forceinline__ void trim_prefix(const unsigned int* __restrict__ hist, int upto, unsigned int k0, int lane,
                                            unsigned int& prefix, unsigned int& krem) {
    prefix = 0;
    krem = k0;
    for (int p = 0; p < upto; p++) {
        int bin = warp_select_bin(hist + p * 256, krem, lane);
        prefix |= (unsigned int)bin << (24 - 8 * p);
    }
}

__global__ void __launch_bounds__(256) trim_hist_kernel(RunConfig cfg, IterState* __restrict__ state,
                                                         const float* __restrict__ distf, int n, unsigned int* __restrict__ hist,
                                                         int pass) {
    if (state->done) return;
    if (pass == 0 && blockIdx.x == 0 && threadIdx.x == 0) stamp_correspondence_end(cfg, state);
    __shared__ unsigned int sh[256];
    __shared__ unsigned int s_prefix;
    sh[threadIdx.x] = 0;
    if (threadIdx.x < 32) {
        unsigned int prefix, krem;
        trim_prefix(hist, pass, (unsigned int)(cfg.n_keep_target - 1), threadIdx.x, prefix, krem);
        if (threadIdx.x == 0) s_prefix = prefix;
    }
    __syncthreads();
    const unsigned int prefix = s_prefix;
    const int shift = 24 - 8 * pass;
    const unsigned int himask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        unsigned int key = trim_key(distf[i], cfg.keep_largest);
        if ((key & himask) == prefix) atomicAdd(&sh[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    unsigned int v = sh[threadIdx.x];
    if (v) atomicAdd(&hist[pass * 256 + threadIdx.x], v);
}

// chunked, order-preserving: block b owns elements [b*chunk, (b+1)*chunk)
__global__ void __launch_bounds__(256) trim_count_eq_kernel(RunConfig cfg, const IterState* __restrict__ state,
                                                             const float* __restrict__ distf, int n,
                                                             const unsigned int* __restrict__ hist, int* __restrict__ block_eq,
                                                             int* __restrict__ eq_total) {
    if (state->done) return;
    __shared__ unsigned int s_thr;
    __shared__ int s_cnt[8];
    if (threadIdx.x < 32) {
        unsigned int prefix, krem;
        trim_prefix(hist, 4, (unsigned int)(cfg.n_keep_target - 1), threadIdx.x, prefix, krem);
        if (threadIdx.x == 0) s_thr = prefix;
    }
    __syncthreads();
    const unsigned int thr = s_thr;
    int chunk = (n + gridDim.x - 1) / gridDim.x;
    int lo = blockIdx.x * chunk, hi = min(n, lo + chunk);
    int c = 0;
    for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) c += trim_key(distf[i], cfg.keep_largest) == thr;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(SE3_FULL, c, o);
    if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int t = 0;
        for (int k = 0; k < 8; k++) t += s_cnt[k];
        block_eq[blockIdx.x] = t;
        if (eq_total && t) atomicAdd(eq_total, t);  // integer sum: order-independent
    }
}

__global__ void __launch_bounds__(256) trim_apply_kernel(RunConfig cfg, IterState* __restrict__ state,
                                                          const float* __restrict__ distf, int n,
                                                          const unsigned int* __restrict__ hist, const int* __restrict__ block_eq,
                                                          const int* __restrict__ rank_eq, int rank, uint8_t* __restrict__ keep) {
    if (state->done) return;
    __shared__ unsigned int s_thr, s_budget;
    __shared__ int s_base;
    __shared__ int s_warp[8];
    if (threadIdx.x < 32) {
        unsigned int prefix, krem;
        trim_prefix(hist, 4, (unsigned int)(cfg.n_keep_target - 1), threadIdx.x, prefix, krem);
        if (threadIdx.x == 0) {
            s_thr = prefix;
            s_budget = krem + 1;  // how many of the elements equal to the threshold survive
            int base = 0;
            if (rank_eq)  // sharded pair: threshold ties owned by lower ranks come first in index order
                for (int r = 0; r < rank; r++) base += rank_eq[r];
            for (int b = 0; b < (int)blockIdx.x; b++) base += block_eq[b];
            s_base = base;
            if (blockIdx.x == 0) {
                state->thr_bits = prefix;
                state->eq_budget = (int)(krem + 1);
                state->n_keep = cfg.n_keep_target;
            }
        }
    }
    __syncthreads();
    const unsigned int thr = s_thr;
    const int budget = (int)s_budget;
    int running = s_base;
    int chunk = (n + gridDim.x - 1) / gridDim.x;
    int lo = blockIdx.x * chunk, hi = min(n, lo + chunk);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int base = lo; base < hi; base += blockDim.x) {
        int i = base + threadIdx.x;
        unsigned int key = i < hi ? trim_key(distf[i], cfg.keep_largest) : 0xffffffffu;
        bool eq = i < hi && key == thr;
        unsigned int m = __ballot_sync(SE3_FULL, eq);
        int in_warp = __popc(m & ((1u << lane) - 1u));
        if (lane == 0) s_warp[w] = __popc(m);
        __syncthreads();
        int before = 0, total = 0;
        for (int k = 0; k < 8; k++) {
            if (k < w) before += s_warp[k];
            total += s_warp[k];
        }
        if (i < hi) keep[i] = (key < thr) || (eq && (running + before + in_warp) < budget);
        running += total;
        __syncthreads();
    }
}

int launch_trim_hist(const RunConfig& cfg, IterState* state, const float* distf, int n, unsigned int* hist, int pass,
                     cudaStream_t st) {
    int g = (n + 255) / 256;
    if (g > 148 * 4) g = 148 * 4;
    if (g < 1) g = 1;
    trim_hist_kernel<<<g, 256, 0, st>>>(cfg, state, distf, n, hist, pass);
    SE3_CUDA(cudaGetLastError());
    return 0;
}

int launch_trim_count_eq(const RunConfig& cfg, IterState* state, const float* distf, int n, const unsigned int* hist,
                         int* block_eq, int* eq_total, cudaStream_t st) {
    trim_count_eq_kernel<<<kReduceBlocks, 256, 0, st>>>(cfg, state, distf, n, hist, block_eq, eq_total);
    SE3_CUDA(cudaGetLastError());
    return 0;
}

int launch_trim_apply(const RunConfig& cfg, IterState* state, const float* distf, int n, const unsigned int* hist,
                      const int* block_eq, const int* rank_eq, int rank, uint8_t* keep, cudaStream_t st) {
    trim_apply_kernel<<<kReduceBlocks, 256, 0, st>>>(cfg, state, distf, n, hist, block_eq, rank_eq, rank, keep);
    SE3_CUDA(cudaGetLastError());
    return 0;
}

int launch_trim(const RunConfig& cfg, IterState* state, CorrBuffers cb, int n, unsigned int* hist, int* block_eq,
                cudaStream_t st) {
    if (!cfg.trim_active) return 0;
    for (int pass = 0; pass < 4; pass++) SE3_TRY(launch_trim_hist(cfg, state, cb.distf, n, hist, pass, st));
    SE3_TRY(launch_trim_count_eq(cfg, state, cb.distf, n, hist, block_eq, nullptr, st));
    SE3_TRY(launch_trim_apply(cfg, state, cb.distf, n, hist, block_eq, nullptr, 0, cb.keep, st));
    return 0;
}

// ------------------------------------------------------------------------------------------------
// single-launch trimmed rejection (single-GPU runs)
// ------------------------------------------------------------------------------------------------
constexpr int kCoarseBlocks = 64;                            // trim_bin_kernel: one block per 1024 fine bins
constexpr int kFinePerCoarse = kTrimHistBins / kCoarseBlocks;  // 1024
constexpr int kSelBlocks = 148;
// CorrBuffers::tcount layout (unsigned int): [0] candidates appended, [1] ticket of trim_bin, [2] ticket of reduce,
// [3] ticket of trim_select, [4 .. 4 + kCoarseBlocks) coarse histogram
static_assert(kTcountWords == 4 + kCoarseBlocks, "tcount layout");

// inclusive scan over the 256 threads of a block (s_warp: 8 words of shared memory)
__device__ __forceinline__ unsigned int block_scan256(unsigned int v, unsigned int* s_warp, unsigned int& total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned int u = __shfl_up_sync(SE3_FULL, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    unsigned int base = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        unsigned int x = s_warp[k];
        if (k < w) base += x;
        tot += x;
    }
    total = tot;
    __syncthreads();
    return base + incl;
}

// Step 1 of 2: which 16-bit bin of the key histogram (filled by the correspondence stage) holds rank n_keep - 1, and
// which rank inside the bin.  64 blocks reduce 1024 bins each to one coarse count; the last block to finish walks the
// 64 coarse counts and then the 1024 fine bins of the one that matters.  Result in IterState::thr_bits (bin) /
// eq_budget (rank in bin); n_keep = -1 flags "fewer counted correspondences than n_keep" (keep everything).
__global__ void __launch_bounds__(256) trim_bin_kernel(RunConfig cfg, IterState* __restrict__ state, CorrBuffers cb) {
    if (state->done) return;
    if (blockIdx.x == 0 && threadIdx.x == 0) stamp_correspondence_end(cfg, state);
    __shared__ unsigned int s_warp[8];
    __shared__ int s_last;
    __shared__ unsigned int s_coarse, s_rank;
    const int tid = threadIdx.x;
    const unsigned int* hist = cb.thist;
    unsigned int* coarse = cb.tcount + 4;
    {
        uint4 v = reinterpret_cast<const uint4*>(hist + blockIdx.x * kFinePerCoarse)[tid];
        unsigned int total;
        block_scan256(v.x + v.y + v.z + v.w, s_warp, total);
        if (tid == 0) coarse[blockIdx.x] = total;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(&cb.tcount[1], 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const unsigned int k0 = (unsigned int)(cfg.n_keep_target - 1);
    {   // coarse level: 64 counts, thread t < 64 owns one
        unsigned int cv = tid < kCoarseBlocks ? __ldcg(coarse + tid) : 0u, total;
        unsigned int incl = block_scan256(cv, s_warp, total);
        if (tid == 0) s_coarse = 0xffffffffu;
        __syncthreads();
        if (tid < kCoarseBlocks && k0 >= incl - cv && k0 < incl) {
            s_coarse = (unsigned int)tid;
            s_rank = k0 - (incl - cv);
        }
        __syncthreads();
    }
    if (s_coarse == 0xffffffffu) {
        if (tid == 0) {
            state->thr_bits = 0xffffffffu;
            state->tie_limit = 0x7fffffff;
            state->n_keep = -1;
            cb.tcount[1] = 0u;
        }
        return;
    }
    {   // fine level: 1024 bins of the chosen coarse bin, 4 consecutive ones per thread
        const unsigned int cbin = s_coarse, k1 = s_rank;
        uint4 v = reinterpret_cast<const uint4*>(hist + cbin * kFinePerCoarse)[tid];
        unsigned int mine = v.x + v.y + v.z + v.w, total;
        unsigned int incl = block_scan256(mine, s_warp, total);
        unsigned int c = incl - mine;
        if (k1 >= c && k1 < incl) {  // exactly one thread
            unsigned int f[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; e++) {
                if (k1 >= c && k1 < c + f[e]) {
                    state->thr_bits = cbin * kFinePerCoarse + (unsigned int)(tid * 4 + e);
                    state->eq_budget = (int)(k1 - c);
                    state->n_keep = cfg.n_keep_target;
                }
                c += f[e];
            }
        }
        if (tid == 0) cb.tcount[1] = 0u;
    }
}

// Step 2 of 2: every block appends the (key, index) pairs of its share of the correspondences that fall into the bin
// and clears its share of the histogram for the next iteration; the last block to finish radix-selects the wanted pair,
// 8 bits at a time from the 16 low key bits down through the index, stopping as soon as one candidate is left.  The
// pair goes to IterState::thr_bits / tie_limit; the reduction keeps correspondence i iff (key_i, i) <= that pair.
__global__ void __launch_bounds__(256) trim_select_kernel(RunConfig cfg, IterState* __restrict__ state, CorrBuffers cb,
                                                           int begin, int end) {
    if (state->done) return;
    __shared__ unsigned int s_hist[256];
    __shared__ unsigned int s_digit, s_rank, s_cnt;
    __shared__ int s_last;
    const int tid = threadIdx.x, lane = tid & 31;
    const unsigned int bin = state->thr_bits;
    const unsigned int rank0 = (unsigned int)state->eq_budget;
    const bool keep_all = state->n_keep < 0;
    for (int b = blockIdx.x * blockDim.x + tid; b < kTrimHistBins; b += gridDim.x * blockDim.x) cb.thist[b] = 0u;
    if (!keep_all) {
        for (int i = begin + blockIdx.x * blockDim.x + tid; i < end; i += gridDim.x * blockDim.x) {
            unsigned int key = trim_key(cb.distf[i], cfg.keep_largest);
            if ((key >> 16) == bin) cb.tcand[atomicAdd(&cb.tcount[0], 1u)] = ((unsigned long long)key << 32) | (unsigned int)i;
        }
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) s_last = atomicAdd(&cb.tcount[3], 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    const unsigned int cnt = keep_all ? 0u : __ldcg(&cb.tcount[0]);
    __syncthreads();  // everyone has read the state and the count before they are rewritten
    if (tid == 0) {
        cb.tcount[0] = 0u;
        cb.tcount[3] = 0u;
    }
    if (keep_all) return;
    unsigned long long prefix = (unsigned long long)bin << 48, mask = 0xffffull << 48;
    unsigned int rank = rank0, remaining = cnt;
    for (int shift = 40; shift >= 0 && remaining > 1; shift -= 8) {
        s_hist[tid] = 0;
        __syncthreads();
        for (unsigned int c = tid; c < cnt; c += blockDim.x) {
            unsigned long long v = __ldcg(&cb.tcand[c]);
            if ((v & mask) == prefix) atomicAdd(&s_hist[(unsigned int)(v >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (tid < 32) {
            unsigned int r = rank;
            int d = warp_select_bin(s_hist, r, lane);
            if (lane == 0) {
                s_digit = (unsigned int)d;
                s_rank = r;
                s_cnt = s_hist[d];
            }
        }
        __syncthreads();
        prefix |= (unsigned long long)s_digit << shift;
        mask |= 0xffull << shift;
        rank = s_rank;
        remaining = s_cnt;
        __syncthreads();
    }
    // with remaining == 1 exactly one pair matches the prefix; otherwise the prefix is the complete pair
    for (unsigned int c = tid; c < cnt; c += blockDim.x) {
        unsigned long long v = __ldcg(&cb.tcand[c]);
        if ((v & mask) == prefix) {
            state->thr_bits = (unsigned int)(v >> 32);
            state->tie_limit = (int)(unsigned int)(v & 0xffffffffull);
            state->eq_budget = (int)rank + 1;
        }
    }
}

int launch_trim_select(const RunConfig& cfg, IterState* state, CorrBuffers cb, int begin, int end, cudaStream_t st) {
    if (!cfg.trim_active || cfg.n_keep_target <= 0) return 0;
    static_assert(kFinePerCoarse == 256 * 4, "trim_bin_kernel: one uint4 per thread");
    trim_bin_kernel<<<kCoarseBlocks, 256, 0, st>>>(cfg, state, cb);
    int g = (end - begin + 1023) / 1024;
    if (g > kSelBlocks) g = kSelBlocks;
    if (g < 1) g = 1;
    trim_select_kernel<<<g, 256, 0, st>>>(cfg, state, cb, begin, end);
    SE3_CUDA(cudaGetLastError());
    return 0;
}

// stage-level API only: histogram of given distances / keep mask from the selected pair
__global__ void __launch_bounds__(256) trim_hist16_kernel(RunConfig cfg, const float* __restrict__ distf, int n,
                                                           unsigned int* __restrict__ hist) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        atomicAdd(&hist[trim_key(distf[i], cfg.keep_largest) >> 16], 1u);
}

__global__ void __launch_bounds__(256) trim_mask_kernel(RunConfig cfg, const IterState* __restrict__ state,
                                                         const float* __restrict__ distf, int n, uint8_t* __restrict__ keep) {
    const unsigned int thr = state->thr_bits;
    const int lim = state->tie_limit;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        keep[i] = trim_keeps(trim_key(distf[i], cfg.keep_largest), i, thr, lim);
}

int launch_trim_stage(const RunConfig& cfg, IterState* state, CorrBuffers cb, int n, cudaStream_t st) {
    int g = (n + 255) / 256;
    if (g > 148 * 4) g = 148 * 4;
    SE3_CUDA(cudaMemsetAsync(cb.thist, 0, kTrimHistBins * sizeof(unsigned int), st));
    SE3_CUDA(cudaMemsetAsync(cb.tcount, 0, kTcountWords * sizeof(unsigned int), st));
    trim_hist16_kernel<<<g, 256, 0, st>>>(cfg, cb.distf, n, cb.thist);
    SE3_TRY(launch_trim_select(cfg, state, cb, 0, n, st));
    trim_mask_kernel<<<g, 256, 0, st>>>(cfg, state, cb.distf, n, cb.keep);
    SE3_CUDA(cudaGetLastError());
    return 0;
}

// ------------------------------------------------------------------------------------------------
// reduction.  Partial record (kReducePartials doubles per block):
//   pt2pl / gicp : [0..20] upper triangle of JTJ (row-major), [21..26] JTr
//   pt2pt        : [0..2] sum(s-a), [3..5] sum(t-b), [6..14] sum (t-b)(s-a)^T row-major
//   all          : [27] sum of distances, [28] count, [29] first-block start time marker (unused)
// ------------------------------------------------------------------------------------------------
constexpr int kAcc = 29;

// ------------------------------------------------------------------------------------------------
// solve + update (one warp)
// ------------------------------------------------------------------------------------------------
// Pivoted LDL^T of the symmetric 6x6 normal matrix (diagonal pivoting: the largest remaining |diagonal| is moved to
// position k, like Eigen::LDLT, which Open3D's SolveLinearSystemPSD calls), then the two triangular solves; a zero
// pivot zeroes its column instead of dividing (rank-deficient systems), a non-finite solution is reported as failure.
// Carried out by a whole thread block on shared memory: thread 6 r + c owns entry (r, c), every
// elimination step is a handful of parallel phases instead of a chain of dependent local-memory accesses in one
// thread (the round-1 single-thread form cost 24 us per iteration; its loops could not be force-unrolled into
// registers because nvcc 12.9 miscompiles that for sm_100a).  tot: 21 upper-triangular JTJ entries (row-major) + 6 JTr; solves JTJ x = -JTr.
// Every thread of the block must call it; x / ok are valid for every thread afterwards.
struct Ldlt6Shared {
    double A[6][6], L[6][6], D[6], y[6], z[6], x[6];
    int perm[6], piv, ok;
};

__device__ void ldlt6_solve_block(const double* tot, Ldlt6Shared& S) {
    const int t = threadIdx.x, r = t / 6, c = t % 6;
    const bool mine = t < 36;
    if (mine) {
        const int a = r < c ? r : c, b = r < c ? c : r;
        S.A[r][c] = tot[a * 6 - a * (a - 1) / 2 + (b - a)];  // row-major upper triangle: row a starts at 6a - a(a-1)/2
        S.L[r][c] = 0.0;
    }
    if (t < 6) S.perm[t] = t;
    __syncthreads();
    for (int k = 0; k < 6; k++) {
        if (t == 0) {
            int piv = k;
            double big = fabs(S.A[k][k]);
            for (int i = k + 1; i < 6; i++) {
                double v = fabs(S.A[i][i]);
                if (v > big) {
                    big = v;
                    piv = i;
                }
            }
            S.piv = piv;
        }
        __syncthreads();
        const int piv = S.piv;
        if (piv != k) {  // symmetric row/column swap k <-> piv of A, row swap of L, entry swap of perm
            const int mr = r == k ? piv : (r == piv ? k : r), mc = c == k ? piv : (c == piv ? k : c);
            double va = 0.0, vl = 0.0;
            int vp = 0;
            if (mine) {
                va = S.A[mr][mc];
                vl = S.L[mr][c];
            }
            if (t < 6) vp = S.perm[t == k ? piv : (t == piv ? k : t)];
            __syncthreads();
            if (mine) {
                S.A[r][c] = va;
                S.L[r][c] = vl;
            }
            if (t < 6) S.perm[t] = vp;
            __syncthreads();
        }
        const double d = S.A[k][k];
        const bool nz = d != 0.0;
        if (t < 6) {
            if (t == k) {
                S.D[k] = d;
                S.L[k][k] = 1.0;
            } else if (t > k) {
                S.L[t][k] = nz ? S.A[t][k] / d : 0.0;
            }
        }
        __syncthreads();
        if (mine && r > k && c > k && nz) S.A[r][c] = S.A[r][c] - S.L[r][k] * d * S.L[c][k];
        __syncthreads();
    }
    if (t == 0) {
        for (int i = 0; i < 6; i++) S.y[i] = -tot[21 + S.perm[i]];  // y = P b
        for (int i = 0; i < 6; i++)
            for (int j = 0; j < i; j++) S.y[i] -= S.L[i][j] * S.y[j];
        for (int i = 0; i < 6; i++) S.z[i] = S.D[i] != 0.0 ? S.y[i] / S.D[i] : 0.0;
        for (int i = 5; i >= 0; i--)
            for (int j = i + 1; j < 6; j++) S.z[i] -= S.L[j][i] * S.z[j];
        int ok = 1;
        for (int i = 0; i < 6; i++) {
            S.x[S.perm[i]] = S.z[i];
            ok = ok && isfinite(S.z[i]);
        }
        S.ok = ok;
    }
    __syncthreads();
}

// Open3D TransformVector6dToMatrix4d: R = Rz(x2) Ry(x1) Rx(x0), t = x3..5 (row-major out)
__device__ void vector6_to_T(const double x[6], double Tm[16]) {
    double cx = cos(x[0]), sx = sin(x[0]), cy = cos(x[1]), sy = sin(x[1]), cz = cos(x[2]), sz = sin(x[2]);
    // Rz*Ry*Rx expanded
    Tm[0] = cz * cy;  Tm[1] = cz * sy * sx - sz * cx;  Tm[2] = cz * sy * cx + sz * sx;  Tm[3] = x[3];
    Tm[4] = sz * cy;  Tm[5] = sz * sy * sx + cz * cx;  Tm[6] = sz * sy * cx - cz * sx;  Tm[7] = x[4];
    Tm[8] = -sy;      Tm[9] = cy * sx;                 Tm[10] = cy * cx;                Tm[11] = x[5];
    Tm[12] = 0.0;     Tm[13] = 0.0;                    Tm[14] = 0.0;                    Tm[15] = 1.0;
}

__device__ double det3(const double A[3][3]) {
    return A[0][0] * (A[1][1] * A[2][2] - A[1][2] * A[2][1]) - A[0][1] * (A[1][0] * A[2][2] - A[1][2] * A[2][0]) +
           A[0][2] * (A[1][0] * A[2][1] - A[1][1] * A[2][0]);
}

// Kabsch rotation from sigma = (1/K) sum (t - mu_t)(s - mu_s)^T  (Eigen::umeyama, no scaling)
__device__ void kabsch_rotation(const double Sg[3][3], double R[3][3]) {
    double a6[6];
    {   // Sg^T Sg
        double G[3][3];
        for (int i = 0; i < 3; i++)
            for (int j = 0; j < 3; j++) G[i][j] = Sg[0][i] * Sg[0][j] + Sg[1][i] * Sg[1][j] + Sg[2][i] * Sg[2][j];
        a6[0] = G[0][0], a6[1] = G[0][1], a6[2] = G[0][2], a6[3] = G[1][1], a6[4] = G[1][2], a6[5] = G[2][2];
    }
    double ev[3], Va[3][3];
    eig3_sym(a6, ev, Va);
    double V[3][3], U[3][3], sv[3];
    for (int c = 0; c < 3; c++) {  // descending singular values
        int src = 2 - c;
        sv[c] = sqrt(fmax(ev[src], 0.0));
        for (int r = 0; r < 3; r++) V[r][c] = Va[r][src];
    }
    double u[3][3];  // u[c] = column c
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++) u[c][r] = Sg[r][0] * V[0][c] + Sg[r][1] * V[1][c] + Sg[r][2] * V[2][c];
    auto nrm = [](const double* v) { return sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); };
    auto dot3 = [](const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; };
    auto cross3 = [](const double* a, const double* b, double* o) {
        o[0] = a[1] * b[2] - a[2] * b[1];
        o[1] = a[2] * b[0] - a[0] * b[2];
        o[2] = a[0] * b[1] - a[1] * b[0];
    };
    double n0 = nrm(u[0]);
    if (n0 > 0.0) { for (int r = 0; r < 3; r++) u[0][r] /= n0; } else { u[0][0] = 1; u[0][1] = 0; u[0][2] = 0; }
    double tiny = 1e-14 * fmax(sv[0], 1e-300);
    double d10 = dot3(u[1], u[0]);
    for (int r = 0; r < 3; r++) u[1][r] -= d10 * u[0][r];
    double n1 = nrm(u[1]);
    if (n1 > tiny) {
        for (int r = 0; r < 3; r++) u[1][r] /= n1;
    } else {
        double t[3] = {fabs(u[0][0]) < 0.9 ? 1.0 : 0.0, fabs(u[0][0]) < 0.9 ? 0.0 : 1.0, 0.0};
        cross3(u[0], t, u[1]);
        double nn = nrm(u[1]);
        for (int r = 0; r < 3; r++) u[1][r] /= nn;
    }
    double d20 = dot3(u[2], u[0]);
    for (int r = 0; r < 3; r++) u[2][r] -= d20 * u[0][r];
    double d21 = dot3(u[2], u[1]);
    for (int r = 0; r < 3; r++) u[2][r] -= d21 * u[1][r];
    double n2 = nrm(u[2]);
    if (n2 > tiny) {
        for (int r = 0; r < 3; r++) u[2][r] /= n2;
    } else {
        cross3(u[0], u[1], u[2]);
    }
    for (int c = 0; c < 3; c++)
        for (int r = 0; r < 3; r++) U[r][c] = u[c][r];
    double d = (det3(U) * det3(V) < 0.0) ? -1.0 : 1.0;
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) R[r][c] = U[r][0] * V[c][0] + U[r][1] * V[c][1] + d * U[r][2] * V[c][2];
}

// x y z + normal or covariance of every target point as one 96-byte record (TargetView::rec)
__global__ void __launch_bounds__(256) pack_target_records_kernel(CloudIndex I, const double* __restrict__ nrm,
                                                                   const double* __restrict__ cov, double* __restrict__ rec) {
    const size_t n = (size_t)I.n;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < I.n; j += gridDim.x * blockDim.x) {
        double r[kTargetRecordDoubles];
#pragma unroll
        for (int e = 0; e < kTargetRecordDoubles; e++) r[e] = 0.0;
        r[0] = I.x[j], r[1] = I.y[j], r[2] = I.z[j];
        if (cov) {
#pragma unroll
            for (int e = 0; e < 6; e++) r[3 + e] = cov[e * n + j];
        } else if (nrm) {
#pragma unroll
            for (int e = 0; e < 3; e++) r[3 + e] = nrm[e * n + j];
        }
        double2* o = reinterpret_cast<double2*>(rec + (size_t)j * kTargetRecordDoubles);
#pragma unroll
        for (int e = 0; e < kTargetRecordDoubles / 2; e++) o[e] = make_double2(r[2 * e], r[2 * e + 1]);
    }
}

int launch_pack_target_records(const CloudIndex& I, const double* nrm, const double* cov, double* rec, cudaStream_t st) {
    int g = (I.n + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    pack_target_records_kernel<<<g, 256, 0, st>>>(I, nrm, cov, rec);
    SE3_CUDA(cudaGetLastError());
    return 0;
}

// fixed-order sum of the per-block records into one record (input of the cross-rank all-reduce)
__global__ void __launch_bounds__(32) sum_partials_kernel(const double* __restrict__ partials, double* __restrict__ total) {
    int lane = threadIdx.x;
    double s = 0.0;
    if (lane < kAcc)
        for (int b = 0; b < kReduceBlocks; b++) s += partials[b * kReducePartials + lane];
    total[lane] = s;
}

int launch_sum_partials(const double* partials, double* total, cudaStream_t st) {
    sum_partials_kernel<<<1, 32, 0, st>>>(partials, total);
    SE3_CUDA(cudaGetLastError());
    return 0;
}

// Executed by one whole block (128 or 256 threads): the stand-alone kernel below (sharded pair, stage API) or the last block
// of reduce_kernel to finish (single-GPU loop).  cond_handle != 0: the caller is the last node of the captured loop
// body and tells the WHILE node whether to run it again.
constexpr unsigned long long kPeerTimeoutNs = 20ull * 1000ull * 1000ull * 1000ull;

// All-reduce (sum) of tot[0 .. kAcc) over the ranks of a sharded pair, through the peer-mapped mailboxes (PeerReduce in
// internal.h).  Called by all 256 threads of one block per rank; returns false when a peer's record did not arrive.
__device__ bool peer_allreduce_block(const PeerReduce& pr, unsigned long long seq, double* tot /*shared*/, int* s_flag /*shared*/) {
    const int tid = threadIdx.x, world = pr.world;
    const int parity = (int)(seq & 1ull);
    const size_t my_slot = (size_t)(parity * world + pr.rank) * kPeerSlotWords;
    // (1) my record into slot [parity][rank] of every mailbox (my own included), 8-byte stores over NVLink
    for (int e = tid; e < world * (kPeerSlotWords - 1); e += blockDim.x) {
        const int peer = e / (kPeerSlotWords - 1), k = e % (kPeerSlotWords - 1);
        const double v = k < kAcc ? tot[k] : 0.0;
        pr.mailboxes[peer][my_slot + k] = (unsigned long long)__double_as_longlong(v);
    }
    __threadfence_system();
    __syncthreads();
    // (2) publish: the sequence word goes last
    if (tid < world) {
        __threadfence_system();  // cumulative: everything the block wrote before the barrier is ordered before the flag
        volatile unsigned long long* flag = pr.mailboxes[tid] + my_slot + (kPeerSlotWords - 1);
        *flag = seq;
    }
    if (tid == 0) *s_flag = 1;
    __syncthreads();
    // (3) wait for the world's records of this iteration in MY mailbox
    const unsigned long long* mine = pr.mailboxes[pr.rank];
    if (tid < world) {
        const volatile unsigned long long* flag = mine + (size_t)(parity * world + tid) * kPeerSlotWords + (kPeerSlotWords - 1);
        const unsigned long long t0 = global_timer_ns();
        while (*flag != seq) {
            if (global_timer_ns() - t0 > kPeerTimeoutNs) {
                *s_flag = 0;
                break;
            }
            __nanosleep(64);
        }
    }
    __threadfence_system();
    __syncthreads();
    if (!*s_flag) return false;
    // (4) sum in rank order: the same operations in the same order on every rank
    if (tid < kAcc) {
        double s = 0.0;
        for (int r = 0; r < world; r++) {
            const volatile unsigned long long* rec = mine + (size_t)(parity * world + r) * kPeerSlotWords;
            s += __longlong_as_double((long long)rec[tid]);
        }
        tot[tid] = s;
    }
    __syncthreads();
    return true;
}

__device__ __noinline__ void solve_update_block(const RunConfig& cfg, IterState* __restrict__ gst,
                                                const double* __restrict__ partials, int n_records,
                                                double* __restrict__ history, unsigned int* __restrict__ hist,
                                                unsigned long long cond_handle, const PeerReduce& peer) {
    // The bookkeeping below touches the state ~150 times from one thread; work on a shared-memory copy (one
    // parallel load, one parallel store) instead of a chain of dependent global accesses.
    static_assert(sizeof(IterState) % 8 == 0, "IterState is copied as 64-bit words");
    __shared__ IterState s_state;
    {
        // L2 loads: when this runs as the tail of reduce_kernel, other blocks have just written parts of the state
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(gst);
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(&s_state);
        for (int k = threadIdx.x; k < (int)(sizeof(IterState) / 8); k += blockDim.x) dst[k] = __ldcg(src + k);
    }
    IterState* st = &s_state;
    __shared__ double tot[kReducePartials];
    __shared__ double wsum[8][kReducePartials];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;  // 4 or 8 warps
    {   // fixed-order sum of the per-block records: warp w takes records w, w + nw, ...; then the warp sums in order
        // (kTailLoads independent L2 loads in flight, added in the fixed order b = w, w + nw, ...: one dependent load per
        // addition made this chain 13 us long, eight in flight still took ten round trips for the 296 records)
        constexpr int kTailLoads = 32;
        double s = 0.0;
        if (lane < kAcc) {
            for (int b0 = w; b0 < n_records; b0 += kTailLoads * nw) {
                double v[kTailLoads];
#pragma unroll
                for (int u = 0; u < kTailLoads; u++) {
                    const int b = b0 + nw * u;
                    v[u] = b < n_records ? __ldcg(&partials[b * kReducePartials + lane]) : 0.0;
                }
#pragma unroll
                for (int u = 0; u < kTailLoads; u++)
                    if (b0 + nw * u < n_records) s += v[u];
            }
        }
        wsum[w][lane] = s;
    }
    if (hist)
        for (int k = threadIdx.x; k < 4 * 256; k += blockDim.x) hist[k] = 0;  // ready for the next iteration's trim
    __syncthreads();
    if (threadIdx.x < kAcc) {
        double s = 0.0;
        for (int k = 0; k < nw; k++) s += wsum[k][threadIdx.x];
        tot[threadIdx.x] = s;
    }
    __syncthreads();
    __shared__ int s_peer_ok;
    if (peer.world > 1) {  // sharded pair: the record becomes the sum over the ranks (block-uniform branch)
        const unsigned long long seq = peer.seq_base + (unsigned long long)(unsigned int)(st->iter + 1);
        if (!peer_allreduce_block(peer, seq, tot, &s_peer_ok)) {
            if (threadIdx.x == 0) {  // abort the run on every rank that sees the time-out; the host reports it
                gst->peer_timeout = 1;
                gst->done = 1;
                if (cond_handle) cudaGraphSetConditional((cudaGraphConditionalHandle)cond_handle, 0u);
            }
            return;
        }
    }
    __shared__ Ldlt6Shared s_ldlt;
    const bool gauss_newton = cfg.variant != SE3ICP_PT2PT && tot[28] > 0.0;  // block-uniform
    if (gauss_newton) ldlt6_solve_block(tot, s_ldlt);
    if (threadIdx.x == 0) {
    const double K = tot[28];
    const double mean = tot[27] / K;  // 0/0 -> NaN exactly as the reference
    double Ti[16];
    for (int k = 0; k < 16; k++) Ti[k] = (k % 5 == 0) ? 1.0 : 0.0;

    if (K > 0.0) {
        if (cfg.variant == SE3ICP_PT2PT) {
            const double* Tm = st->T_total;
            double c0[3] = {cfg.has_se3 ? 0.0 : st->c_src[0], cfg.has_se3 ? 0.0 : st->c_src[1],
                            cfg.has_se3 ? 0.0 : st->c_src[2]};
            double a[3], b[3] = {0, 0, 0};
            for (int r = 0; r < 3; r++) a[r] = Tm[4 * r] * c0[0] + Tm[4 * r + 1] * c0[1] + Tm[4 * r + 2] * c0[2] + Tm[4 * r + 3];
            if (!cfg.has_se3)
                for (int r = 0; r < 3; r++) b[r] = st->c_tgt[r];
            double ms[3] = {tot[0] / K, tot[1] / K, tot[2] / K};  // mu_s - a
            double mt[3] = {tot[3] / K, tot[4] / K, tot[5] / K};  // mu_t - b
            double Sg[3][3];
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) Sg[r][c] = tot[6 + 3 * r + c] / K - mt[r] * ms[c];
            double R[3][3];
            kabsch_rotation(Sg, R);
            double mus[3] = {ms[0] + a[0], ms[1] + a[1], ms[2] + a[2]};
            double mut[3] = {mt[0] + b[0], mt[1] + b[1], mt[2] + b[2]};
            for (int r = 0; r < 3; r++) {
                for (int c = 0; c < 3; c++) Ti[4 * r + c] = R[r][c];
                Ti[4 * r + 3] = mut[r] - (R[r][0] * mus[0] + R[r][1] * mus[1] + R[r][2] * mus[2]);
            }
        } else {
            if (s_ldlt.ok) vector6_to_T(s_ldlt.x, Ti);
        }
    }

    // bookkeeping: reference .cpp:684-686, 709-711
    st->mse_prev = st->mse_cur;
    st->mse_cur = mean;
    st->mse_rel = fabs(st->mse_cur - st->mse_prev);
    double To[16], Tn[16];  // registers: this thread is a chain of dependent operations, keep shared memory out of it
#pragma unroll
    for (int k = 0; k < 16; k++) To[k] = st->T_total[k];
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
        for (int c = 0; c < 4; c++) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 4; k++) s += Ti[4 * r + k] * To[4 * k + c];
            Tn[4 * r + c] = s;
        }
    double ch = 0.0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        double df = To[k] - Tn[k];
        ch += df * df;
        st->T_prev[k] = To[k];
        st->T_i[k] = Ti[k];
        st->T_total[k] = Tn[k];
    }
    st->T_change = sqrt(ch);
    if (cfg.record_history && history && st->hist_count < cfg.max_history) {
        for (int k = 0; k < 16; k++) history[16 * (size_t)st->hist_count + k] = Ti[k];
        st->hist_count++;
    }

    const bool se3_now = se3_phase_active_o(cfg, st);
    st->iter += 1;
    if (se3_now) st->se3_iters += 1;
    const double s = st->scale;
    if (!cfg.has_se3) {  // run_icp .cpp:547-550
        if (st->iter == cfg.max_iter || st->mse_rel < cfg.mse) st->done = 1;
    } else if (cfg.pure) {  // run_se3_pure .cpp:1118
        if (st->iter == cfg.max_se3_iter || st->mse_rel < s * cfg.mse) st->done = 1;
    } else if (!st->switch_icp) {  // .cpp:718-723
        if (st->iter == cfg.max_se3_iter || st->T_change < cfg.mse_switch) {
            st->switch_icp = 1;
            st->switch_iter = st->iter;
            st->t_switch = global_timer_ns();
        }
    } else {  // .cpp:724-729
        if (st->iter == cfg.max_iter || st->mse_rel < s * cfg.mse) st->done = 1;
    }
    if (st->iter >= 1000000) st->done = 1;  // hard cap (the reference would spin forever on such parameters)
    st->total_repairs += st->repair_count;
    st->repair_count = 0;
    st->searched_total += (unsigned long long)st->work_count;
    st->work_count = 0;
    st->corr_stamped = 0;
    st->t_mark = global_timer_ns();
    if (cond_handle) cudaGraphSetConditional((cudaGraphConditionalHandle)cond_handle, st->done ? 0u : 1u);
    }  // thread 0
    __syncthreads();
    {
        const unsigned long long* src = reinterpret_cast<const unsigned long long*>(&s_state);
        unsigned long long* dst = reinterpret_cast<unsigned long long*>(gst);
        for (int k = threadIdx.x; k < (int)(sizeof(IterState) / 8); k += blockDim.x) dst[k] = src[k];
    }
}

__global__ void __launch_bounds__(256) solve_update_kernel(RunConfig cfg, IterState* __restrict__ gst,
                                                            const double* __restrict__ partials, int n_records,
                                                            double* __restrict__ history, unsigned int* __restrict__ hist,
                                                            unsigned long long cond_handle) {
    if (gst->done) {
        if (cond_handle && threadIdx.x == 0) cudaGraphSetConditional((cudaGraphConditionalHandle)cond_handle, 0u);
        return;
    }
    solve_update_block(cfg, gst, partials, n_records, history, hist, cond_handle, PeerReduce{nullptr, 1, 0, 0ull});
}

// fuse.enabled: the block that finishes last also sums the per-block records (fixed order, so the result does not depend
// on which block that is), solves, updates the estimate and decides whether the loop goes on — an iteration then ends
// with this kernel.
constexpr int kReduceThreads = 128;  // 168 registers per thread: three 128-thread blocks fit an SM, so the 296 blocks run in one
                                     // wave (256-thread blocks: one per SM, two waves)
__global__ void __launch_bounds__(kReduceThreads) reduce_kernel(SourceView S, TargetView T, RunConfig cfg,
                                                                 IterState* __restrict__ state, CorrBuffers cb,
                                                                 double* __restrict__ partials, SolveFusion fuse) {
    if (state->done) {
        if (fuse.enabled && fuse.cond_handle && blockIdx.x == 0 && threadIdx.x == 0)
            cudaGraphSetConditional((cudaGraphConditionalHandle)fuse.cond_handle, 0u);
        return;
    }
    __shared__ double Tm[16];
    __shared__ double sm[kReduceThreads / 32][kAcc];
    __shared__ int s_last;
    if (threadIdx.x < 16) Tm[threadIdx.x] = state->T_total[threadIdx.x];
    if (blockIdx.x == 0 && threadIdx.x == 0) stamp_correspondence_end(cfg, state);
    // trimmed rejection: either the (key, index) threshold of trim_select_kernel or the mask of the multi-pass kernels
    const bool trim_thr = cfg.trim_active && cb.thist != nullptr;
    const unsigned int thr_bits = trim_thr ? state->thr_bits : 0u;
    const int tie_limit = trim_thr ? state->tie_limit : 0;
    __syncthreads();
    double acc[kAcc];
#pragma unroll
    for (int k = 0; k < kAcc; k++) acc[k] = 0.0;

    const size_t n = (size_t)S.n, m = (size_t)T.n;
    // reference points that keep the pt2pt sums well conditioned (exact algebra for any a, b)
    double ax = 0, ay = 0, az = 0, bx = 0, by = 0, bz = 0;
    if (cfg.variant == SE3ICP_PT2PT) {
        double c0x = cfg.has_se3 ? 0.0 : state->c_src[0], c0y = cfg.has_se3 ? 0.0 : state->c_src[1],
               c0z = cfg.has_se3 ? 0.0 : state->c_src[2];
        ax = Tm[0] * c0x + Tm[1] * c0y + Tm[2] * c0z + Tm[3];
        ay = Tm[4] * c0x + Tm[5] * c0y + Tm[6] * c0z + Tm[7];
        az = Tm[8] * c0x + Tm[9] * c0y + Tm[10] * c0z + Tm[11];
        if (!cfg.has_se3) {
            bx = state->c_tgt[0];
            by = state->c_tgt[1];
            bz = state->c_tgt[2];
        }
    }

    for (int i = S.begin + blockIdx.x * blockDim.x + threadIdx.x; i < S.end; i += gridDim.x * blockDim.x) {
        if (cfg.trim_active) {
            if (cfg.n_keep_target <= 0) continue;
            if (trim_thr ? !trim_keeps(trim_key(cb.distf[i], cfg.keep_largest), i, thr_bits, tie_limit) : !cb.keep[i]) continue;
        }
        const int j = cb.idx[i];
        if (j < 0) continue;
        const double px = S.x[i], py = S.y[i], pz = S.z[i];
        const double sx = Tm[0] * px + Tm[1] * py + Tm[2] * pz + Tm[3];
        const double sy = Tm[4] * px + Tm[5] * py + Tm[6] * pz + Tm[7];
        const double sz = Tm[8] * px + Tm[9] * py + Tm[10] * pz + Tm[11];
        const double2* __restrict__ trec = T.rec ? reinterpret_cast<const double2*>(T.rec + (size_t)j * kTargetRecordDoubles) : nullptr;
        double tx, ty, tz, t3 = 0.0;  // t3: first entry of the normal / covariance, shares a 16-byte load with z
        if (trec) {
            const double2 a = trec[0], b = trec[1];
            tx = a.x, ty = a.y, tz = b.x, t3 = b.y;
        } else {
            tx = T.idx.x[j], ty = T.idx.y[j], tz = T.idx.z[j];
        }
        const double dx = sx - tx, dy = sy - ty, dz = sz - tz;
        acc[28] += 1.0;
        if (cfg.with_cf)
            acc[27] += sqrt(dx * dx + dy * dy + dz * dz);  // .cpp:396 recomputed in double
        else
            acc[27] += (double)cb.distf[i];  // .cpp:383 stored float distance

        if (cfg.variant == SE3ICP_PT2PT) {
            double us = sx - ax, vs = sy - ay, ws = sz - az;
            double ut = tx - bx, vt = ty - by, wt = tz - bz;
            acc[0] += us, acc[1] += vs, acc[2] += ws;
            acc[3] += ut, acc[4] += vt, acc[5] += wt;
            acc[6] += ut * us, acc[7] += ut * vs, acc[8] += ut * ws;
            acc[9] += vt * us, acc[10] += vt * vs, acc[11] += vt * ws;
            acc[12] += wt * us, acc[13] += wt * vs, acc[14] += wt * ws;
        } else if (cfg.variant == SE3ICP_PT2PL) {
            double nx, ny, nz;
            if (trec) {
                const double2 c = trec[2];
                nx = t3, ny = c.x, nz = c.y;
            } else {
                nx = T.nrm[j], ny = T.nrm[m + j], nz = T.nrm[2 * m + j];
            }
            const double r = dx * nx + dy * ny + dz * nz;
            double J[6] = {sy * nz - sz * ny, sz * nx - sx * nz, sx * ny - sy * nx, nx, ny, nz};
            int o = 0;
#pragma unroll
            for (int a = 0; a < 6; a++) {
#pragma unroll
                for (int b = a; b < 6; b++) acc[o++] += J[a] * J[b];
            }
#pragma unroll
            for (int a = 0; a < 6; a++) acc[21 + a] += J[a] * r;
        } else {
            // M = Ct + R Cs0 R^T ; B = M^-1 (times w^2) ; A = [-[s]x | I] ; JTJ += A^T B A ; JTr += A^T B d
            double c0[6];
#pragma unroll
            for (int e = 0; e < 6; e++) c0[e] = S.cov[e * n + i];
            double R[3][3] = {{Tm[0], Tm[1], Tm[2]}, {Tm[4], Tm[5], Tm[6]}, {Tm[8], Tm[9], Tm[10]}};
            double C[3][3] = {{c0[0], c0[1], c0[2]}, {c0[1], c0[3], c0[4]}, {c0[2], c0[4], c0[5]}};
            double RC[3][3];
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
                for (int c = 0; c < 3; c++) RC[r][c] = R[r][0] * C[0][c] + R[r][1] * C[1][c] + R[r][2] * C[2][c];
            double ct[6];  // target covariance
            if (trec) {
                const double2 c = trec[2], d = trec[3], f = trec[4];
                ct[0] = t3, ct[1] = c.x, ct[2] = c.y, ct[3] = d.x, ct[4] = d.y, ct[5] = f.x;
            } else {
#pragma unroll
                for (int e = 0; e < 6; e++) ct[e] = T.cov[e * m + j];
            }
            double Mm[6];
            {
                int e = 0;
#pragma unroll
                for (int r = 0; r < 3; r++)
#pragma unroll
                    for (int c = r; c < 3; c++) {
                        double v = RC[r][0] * R[c][0] + RC[r][1] * R[c][1] + RC[r][2] * R[c][2];
                        Mm[e] = v + ct[e];
                        e++;
                    }
            }
            double Bi[6];
            sym3_inverse(Mm, Bi);
            double w2 = 1.0;
            if (cfg.with_cf) {
                double w = (S.conf[i] + T.conf[j]) / 2.0;  // .cpp:913
                w2 = w * w;
            }
            double B[3][3] = {{Bi[0] * w2, Bi[1] * w2, Bi[2] * w2},
                              {Bi[1] * w2, Bi[3] * w2, Bi[4] * w2},
                              {Bi[2] * w2, Bi[4] * w2, Bi[5] * w2}};
            // columns of A = [-[s]x | I]: a0 = e_x x s, a1 = e_y x s, a2 = e_z x s, a3..5 = e_x, e_y, e_z
            double A[6][3] = {{0.0, -sz, sy}, {sz, 0.0, -sx}, {-sy, sx, 0.0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
            double BA[6][3];
#pragma unroll
            for (int a = 0; a < 6; a++)
#pragma unroll
                for (int r = 0; r < 3; r++) BA[a][r] = B[r][0] * A[a][0] + B[r][1] * A[a][1] + B[r][2] * A[a][2];
            int o = 0;
#pragma unroll
            for (int a = 0; a < 6; a++) {
#pragma unroll
                for (int b = a; b < 6; b++) acc[o++] += A[a][0] * BA[b][0] + A[a][1] * BA[b][1] + A[a][2] * BA[b][2];
            }
#pragma unroll
            for (int a = 0; a < 6; a++) acc[21 + a] += BA[a][0] * dx + BA[a][1] * dy + BA[a][2] * dz;
        }
    }

    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < kAcc; k++) {
        double v = warp_sum(acc[k]);
        if (lane == 0) sm[w][k] = v;
    }
    __syncthreads();
    if (threadIdx.x < kAcc) {
        double s = 0.0;
        for (int k = 0; k < kReduceThreads / 32; k++) s += sm[k][threadIdx.x];
        partials[blockIdx.x * kReducePartials + threadIdx.x] = s;
    }
    if (!fuse.enabled) return;
    // last block to arrive finishes the iteration
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&cb.tcount[2], 1u) == gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    __threadfence();
    if (threadIdx.x == 0) cb.tcount[2] = 0u;  // ready for the next launch
    solve_update_block(cfg, state, partials, (int)gridDim.x, fuse.history, fuse.hist, fuse.cond_handle, fuse.peer);
}

int launch_reduce(const SourceView& S, const TargetView& T, const RunConfig& cfg, IterState* state, CorrBuffers cb,
                  double* partials, const SolveFusion& fuse, cudaStream_t st) {
    reduce_kernel<<<kReduceBlocks, kReduceThreads, 0, st>>>(S, T, cfg, state, cb, partials, fuse);
    SE3_CUDA(cudaGetLastError());
    return 0;
}

int launch_solve_update(const RunConfig& cfg, IterState* state, const double* partials, int n_records, double* history,
                        unsigned int* hist, unsigned long long cond_handle, cudaStream_t st) {
    solve_update_kernel<<<1, 256, 0, st>>>(cfg, state, partials, n_records, history, hist, cond_handle);
    SE3_CUDA(cudaGetLastError());
    return 0;
}

// reference .cpp:735-738: t = t'/s - R' c_src + c_tgt (rotation unchanged)
__global__ void finalize_kernel(RunConfig cfg, IterState* st) {
    if (threadIdx.x != 0) return;
    for (int k = 0; k < 16; k++) st->T_final[k] = st->T_total[k];
    if (cfg.has_se3) {
        double inv = 1.0 / st->scale;
        for (int r = 0; r < 3; r++) {
            double rc = st->T_total[4 * r] * st->c_src[0] + st->T_total[4 * r + 1] * st->c_src[1] +
                        st->T_total[4 * r + 2] * st->c_src[2];
            st->T_final[4 * r + 3] = inv * st->T_total[4 * r + 3] - rc + st->c_tgt[r];
        }
    }
}

// last node of the captured iteration: tells the WHILE node of the CUDA graph whether to run the body again
__global__ void loop_condition_kernel(cudaGraphConditionalHandle handle, const IterState* __restrict__ st) {
    if (threadIdx.x == 0) cudaGraphSetConditional(handle, st->done ? 0u : 1u);
}

int launch_loop_condition(unsigned long long handle, const IterState* state, cudaStream_t st) {
    loop_condition_kernel<<<1, 32, 0, st>>>((cudaGraphConditionalHandle)handle, state);
    SE3_CUDA(cudaGetLastError());
    return 0;
}

int launch_finalize(const RunConfig& cfg, IterState* state, cudaStream_t st) {
    finalize_kernel<<<1, 32, 0, st>>>(cfg, state);
    SE3_CUDA(cudaGetLastError());
    return 0;
}

// end of the set-up stage: the correspondence timers of the first iteration start here, not at the start of the run
__global__ void mark_loop_start_kernel(IterState* st) {
    if (threadIdx.x == 0) st->t_mark = global_timer_ns();
}

int launch_mark_loop_start(IterState* state, cudaStream_t st) {
    mark_loop_start_kernel<<<1, 32, 0, st>>>(state);
    SE3_CUDA(cudaGetLastError());
    return 0;
}

int launch_init_state(IterState* state, unsigned int* hist, cudaStream_t st) {
    init_state_kernel<<<1, 32, 0, st>>>(state, hist);
    SE3_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace se3
