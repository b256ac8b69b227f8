// capi.cu — C ABI (include/se3icp.h): context management, the registration driver that enqueues the
// kernels of spatial_index.cu / knn_features.cu / nn_search.cu / optimise.cu, the batch runner and
// the stage-level entry points used by the parity tests.  No CPU fallback anywhere: every entry
// point fails with a status code when CUDA fails.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cmath>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"
#include "context.h"
#include "internal.h"
#include "nccl_dyn.h"

namespace se3 {

static thread_local char g_err[1024] = "";

void set_last_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int DeviceBuf::ensure(size_t bytes) {
    if (bytes <= cap && ptr) return 0;
    release();
    size_t want = bytes < 256 ? 256 : bytes;
    cudaError_t e = cudaMalloc(&ptr, want);
    if (e != cudaSuccess) {
        ptr = nullptr;
        cap = 0;
        set_last_error("cudaMalloc(%zu) -> %s", want, cudaGetErrorString(e));
        return SE3ICP_ERR_CUDA;
    }
    cap = want;
    return 0;
}

int DeviceBuf::ensure_keep(size_t bytes, size_t keep, cudaStream_t st) {
    if (bytes <= cap && ptr) return 0;
    void* np = nullptr;
    size_t want = bytes + bytes / 2;
    cudaError_t e = cudaMalloc(&np, want);
    if (e != cudaSuccess) {
        set_last_error("cudaMalloc(%zu) -> %s", want, cudaGetErrorString(e));
        return SE3ICP_ERR_CUDA;
    }
    if (ptr && keep) {
        e = cudaMemcpyAsync(np, ptr, keep, cudaMemcpyDeviceToDevice, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) {
            cudaFree(np);
            set_last_error("growing a device buffer: copy of %zu bytes -> %s", keep, cudaGetErrorString(e));
            return SE3ICP_ERR_CUDA;
        }
    }
    release();
    ptr = np;
    cap = want;
    return 0;
}

void DeviceBuf::release() {
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    cap = 0;
}

}  // namespace se3

using namespace se3;

SourceView se3icp_ctx::source_view() const {
    SourceView S;
    S.n = (int)n[0];
    S.begin = sharded ? shard_begin : 0;
    S.end = sharded ? shard_end : (int)n[0];
    static const bool use_order = [] {
        const char* e = getenv("SE3ICP_SRC_ORDER");
        return !(e && atoi(e) == 0);
    }();
    S.order = (use_order && !sharded && src_index_built) ? index[0].perm.as<int>() : nullptr;
    S.x = index[0].x.as<double>();
    S.y = index[0].y.as<double>();
    S.z = index[0].z.as<double>();
    S.frame = frame[0].as<double>();
    S.cov = cov[0].as<double>();
    S.conf = conf[0].as<double>();
    return S;
}

TargetView se3icp_ctx::target_view() const {
    TargetView T;
    T.n = (int)n[1];
    T.idx = index[1].view;
    T.nrm = nrm[1].as<double>();
    T.cov = cov[1].as<double>();
    T.conf = conf[1].as<double>();
    T.rec = tgt_rec_valid ? tgt_rec.as<double>() : nullptr;
    T.rows32 = se3idx.rows32.as<float4>();
    T.rows64 = se3idx.rows64.as<double>();
    T.box12 = se3idx.box12.as<float2>();
    T.perm12 = se3idx.perm12.as<int>();
    T.inv12 = se3idx.inv12.as<int>();
    T.keys12 = se3idx.keys12.as<uint64_t>();
    T.tscale = cfg.with_cf ? 1.0 : cfg.beta;  // .cpp:834-836: the _with_cf tree holds the unscaled point
    T.dist_scale = T.tscale != 0.0 ? cfg.beta / T.tscale : 0.0;
    return T;
}

CorrBuffers se3icp_ctx::corr_buffers(bool with_d2) const {
    CorrBuffers cb;
    cb.d2_nd = with_d2 ? d2_nd.as<double>() : nullptr;
    cb.idx = corr_idx.as<int>();
    cb.dist = corr_dist.as<double>();
    cb.distf = corr_distf.as<float>();
    cb.keep = keep.as<uint8_t>();
    cb.repair = repair.as<int>();
    cb.work = work.as<int>();
    cb.ref_iter = ref_iter.as<int>();
    cb.t_table = t_table.as<double>();
    cb.t_table_cap = t_table.ptr ? t_table_cap : 0;
    cb.ref_d2nd = ref_d2nd.as<double>();
    // single-launch trimmed rejection whenever no cross-rank histogram exchange is needed
    const bool thr_trim = cfg.trim_active && cfg.n_keep_target > 0 && !sharded;
    cb.thist = thr_trim ? thist.as<unsigned int>() : nullptr;
    cb.tcand = tcand.as<unsigned long long>();
    cb.tcount = tcount.as<unsigned int>();
    return cb;
}

namespace {

// PCL CorrespondenceRejectorTrimmed: overlap stored as float; floor(ratio * float(N)) as unsigned
size_t trimmed_count(size_t n, double overlap) {
    float ratio = (float)overlap;
    float prod = ratio * static_cast<float>(n);
    double fl = std::floor((double)prod);
    if (fl < 0.0) fl = 0.0;
    if (fl > 4294967295.0) fl = 4294967295.0;
    return (size_t)(unsigned int)fl;
}

int fill_config(se3icp_ctx* c, const se3icp_params* p) {
    if (p->entry < SE3ICP_RUN_ICP || p->entry > SE3ICP_RUN_SE3_PURE) {
        set_last_error("bad entry %d", p->entry);
        return SE3ICP_ERR_ARG;
    }
    if (p->variant < SE3ICP_PT2PT || p->variant > SE3ICP_GICP) {
        set_last_error("bad variant %d (valid: pt2pt, pt2pl, gicp)", p->variant);
        return SE3ICP_ERR_ARG;
    }
    RunConfig& cfg = c->cfg;
    cfg.entry = p->entry;
    cfg.with_cf = p->entry == SE3ICP_RUN_SE3_ICP_CF;
    cfg.variant = cfg.with_cf ? (int)SE3ICP_GICP : p->variant;  // .cpp:855-856,921: GICP hard-wired
    cfg.max_iter = p->max_num_iterations;
    cfg.max_se3_iter = p->max_num_se3_iterations;
    cfg.has_se3 = p->entry != SE3ICP_RUN_ICP;
    cfg.pure = p->entry == SE3ICP_RUN_SE3_PURE;
    size_t N = c->n[0];
    size_t keep = trimmed_count(N, p->estimated_overlap);
    cfg.trim_active = keep < N;
    cfg.n_keep_target = (int)std::min(keep, N);
    cfg.keep_largest = p->trim_keep_largest != 0;
    cfg.record_history = p->record_history != 0;
    long cap = std::max<long>(std::max(p->max_num_iterations, p->max_num_se3_iterations), 1);
    cfg.max_history = (int)std::min<long>(cap, 100000);
    cfg.coherence = p->nn_coherence != 0 && cfg.has_se3;
    cfg.coherence_xyz = p->nn_coherence != 0 && !cfg.pure;
    // Tracking the second-nearest distance costs ~2 % of a search (measured), so the filter is always armed;
    // the threshold remains as a tuning knob (||T_prev - T_total||_F of the last iteration).
    cfg.coherence_thr = 1e300;
    // while the estimate still jumps, the remembered match is a poor starting point (tuning knob; exactness unaffected)
    static const double reseed_thr = [] {
        const char* e = getenv("SE3ICP_RESEED_THR");
        return e && *e ? atof(e) : 0.05;
    }();
    cfg.reseed_thr = reseed_thr;
    cfg.mse = p->mse;
    cfg.mse_switch = p->mse_switch_error;
    cfg.alpha = p->alpha_rot;
    cfg.beta = p->beta_transl;
    c->params = *p;
    return 0;
}

int alloc_run(se3icp_ctx* c) {
    const RunConfig& cfg = c->cfg;
    size_t N = c->n[0], M = c->n[1];
    SE3_TRY(c->index[0].reserve((int)N));
    SE3_TRY(c->index[1].reserve((int)M));
    for (int w = 0; w < 2; w++) {
        size_t nn = c->n[w];
        if (cfg.has_se3) SE3_TRY(c->frame[w].ensure(9 * nn * sizeof(double)));
        if (cfg.variant == SE3ICP_GICP || (w == 1 && cfg.variant == SE3ICP_PT2PL))
            SE3_TRY(c->nrm[w].ensure(3 * nn * sizeof(double)));
        if (cfg.variant == SE3ICP_GICP) SE3_TRY(c->cov[w].ensure(6 * nn * sizeof(double)));
        if (cfg.with_cf) SE3_TRY(c->conf[w].ensure(nn * sizeof(double)));
        SE3_TRY(c->psum[w].ensure((size_t)kReduceBlocks * 3 * sizeof(double)));
        SE3_TRY(c->pmax[w].ensure((size_t)kReduceBlocks * sizeof(double)));
    }
    if (cfg.has_se3) SE3_TRY(c->se3idx.reserve((int)M, c->index[1].view));
    SE3_TRY(c->corr_idx.ensure(N * sizeof(int)));
    SE3_TRY(c->corr_dist.ensure(N * sizeof(double)));
    SE3_TRY(c->corr_distf.ensure(N * sizeof(float)));
    SE3_TRY(c->keep.ensure(N));
    SE3_TRY(c->repair.ensure(N * sizeof(int)));
    SE3_TRY(c->d2_nd.ensure(N * sizeof(double)));
    if (cfg.coherence || cfg.coherence_xyz) {
        SE3_TRY(c->ref_iter.ensure(N * sizeof(int)));
        c->t_table_cap = (int)std::min<long>(std::max<long>(std::max(cfg.max_iter, cfg.max_se3_iter), 1) + 2, 100000);
        SE3_TRY(c->t_table.ensure((size_t)c->t_table_cap * 16 * sizeof(double)));
        SE3_TRY(c->ref_d2nd.ensure(N * sizeof(double)));
        SE3_TRY(c->work.ensure(N * sizeof(int)));
    }
    SE3_TRY(c->partials.ensure((size_t)kReduceBlocks * kReducePartials * sizeof(double)));
    SE3_TRY(c->hist.ensure(4 * 256 * sizeof(unsigned int)));
    SE3_TRY(c->block_eq.ensure((size_t)kReduceBlocks * sizeof(int)));
    SE3_TRY(c->tcount.ensure(kTcountWords * sizeof(unsigned int)));
    if (cfg.trim_active) {
        SE3_TRY(c->thist.ensure((size_t)kTrimHistBins * sizeof(unsigned int)));
        SE3_TRY(c->tcand.ensure(N * sizeof(unsigned long long)));
    }
    if (cfg.record_history) SE3_TRY(c->history.ensure((size_t)cfg.max_history * 16 * sizeof(double)));
    SE3_TRY(c->state.ensure(sizeof(IterState)));
    SE3_TRY(c->totals.ensure(kReducePartials * sizeof(double)));
    SE3_TRY(c->eq_total.ensure(sizeof(int)));
    SE3_TRY(c->rank_eq.ensure((size_t)std::max(c->comm_size, 1) * sizeof(int)));
    return 0;
}

#define SE3_NCCL(call)                                                                        \
    do {                                                                                      \
        ncclResult_t r__ = (call);                                                            \
        if (r__ != ncclSuccess) {                                                             \
            set_last_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, nccl->GetErrorString(r__)); \
            return SE3ICP_ERR_NCCL;                                                           \
        }                                                                                     \
    } while (0)

// ---- peer-memory mailboxes of the sharded pair -------------------------------------------------------------------
void peer_teardown(se3icp_ctx* c) {
    for (int r = 0; r < kMaxPeers; r++) {
        if (c->peer_ptr[r] && r != c->comm_rank) cudaIpcCloseMemHandle(c->peer_ptr[r]);
        c->peer_ptr[r] = nullptr;
    }
    c->peer_ready = false;
}

// Every rank allocates its mailbox, the CUDA IPC handles travel through one ncclAllGather on the communicator the
// context already holds, and every rank maps the others' mailboxes (NVLink peer access).  Not fatal when it fails
// (GPUs without peer access, ranks on different nodes): the run then all-reduces through NCCL from a host-driven loop.
int peer_setup(se3icp_ctx* c) {
    peer_teardown(c);
    static const bool disabled = [] {
        const char* e = getenv("SE3ICP_SHARDED_P2P");
        return e && atoi(e) == 0;
    }();
    const int world = c->comm_size;
    if (disabled || !c->comm || world < 2 || world > kMaxPeers) return 0;
    const NcclApi* nccl = nccl_api();
    if (!nccl) return 0;
    cudaStream_t st = c->stream;
    const size_t box_bytes = (size_t)2 * world * kPeerSlotWords * sizeof(unsigned long long);
    SE3_TRY(c->mailbox.ensure(box_bytes));
    SE3_TRY(c->mailbox_table.ensure((size_t)kMaxPeers * sizeof(void*)));
    SE3_TRY(c->scratch.ensure((size_t)(world + 1) * sizeof(cudaIpcMemHandle_t)));
    SE3_CUDA(cudaMemsetAsync(c->mailbox.ptr, 0, c->mailbox.cap, st));
    cudaIpcMemHandle_t mine;
    SE3_CUDA(cudaIpcGetMemHandle(&mine, c->mailbox.ptr));
    std::vector<cudaIpcMemHandle_t> all(world);
    char* d_mine = c->scratch.as<char>();
    char* d_all = d_mine + sizeof(cudaIpcMemHandle_t);
    SE3_CUDA(cudaMemcpyAsync(d_mine, &mine, sizeof(mine), cudaMemcpyHostToDevice, st));
    SE3_NCCL(nccl->AllGather(d_mine, d_all, sizeof(mine), ncclChar, (ncclComm_t)c->comm, st));
    SE3_CUDA(cudaMemcpyAsync(all.data(), d_all, (size_t)world * sizeof(mine), cudaMemcpyDeviceToHost, st));
    SE3_CUDA(cudaStreamSynchronize(st));  // also: every rank's mailbox is zeroed before anyone can write into it ...
    bool ok = true;
    for (int r = 0; r < world && ok; r++) {
        if (r == c->comm_rank) {
            c->peer_ptr[r] = c->mailbox.ptr;
        } else if (cudaIpcOpenMemHandle(&c->peer_ptr[r], all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            c->peer_ptr[r] = nullptr;
            ok = false;
        }
    }
    // ... which the second collective guarantees: nobody leaves it before everybody has passed the synchronize above.
    // It also agrees on the outcome: one rank without peer access sends everybody to the NCCL path.
    int* d_ok = reinterpret_cast<int*>(d_mine);
    int h_ok = ok ? 1 : 0;
    SE3_CUDA(cudaMemcpyAsync(d_ok, &h_ok, sizeof(int), cudaMemcpyHostToDevice, st));
    SE3_NCCL(nccl->AllReduce(d_ok, d_ok, 1, ncclInt32, ncclMin, (ncclComm_t)c->comm, st));
    SE3_CUDA(cudaMemcpyAsync(&h_ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, st));
    SE3_CUDA(cudaStreamSynchronize(st));
    if (!h_ok) {
        peer_teardown(c);
        return 0;
    }
    SE3_CUDA(cudaMemcpyAsync(c->mailbox_table.ptr, c->peer_ptr, (size_t)kMaxPeers * sizeof(void*), cudaMemcpyHostToDevice, st));
    SE3_CUDA(cudaStreamSynchronize(st));
    c->peer_ready = true;
    c->peer_runs = 0;
    return 0;
}

// SE3ICP_SETUP_TIMING=1: CUDA events between the stages of the set-up, printed on stderr by se3icp_run_finish
struct SetupMarks {
    bool on = false;
    std::vector<std::pair<const char*, cudaEvent_t>> marks;
    SetupMarks() {
        const char* e = getenv("SE3ICP_SETUP_TIMING");
        on = e && atoi(e) != 0;
    }
    void mark(const char* name, cudaStream_t st) {
        if (!on) return;
        cudaEvent_t ev;
        if (cudaEventCreate(&ev) != cudaSuccess) return;
        cudaEventRecord(ev, st);
        marks.emplace_back(name, ev);
    }
    void report(int rank) {
        if (!on || marks.size() < 2) return;
        std::string line = "[se3icp] set-up stages (rank " + std::to_string(rank) + "):";
        for (size_t k = 1; k < marks.size(); k++) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, marks[k - 1].second, marks[k].second);
            char buf[96];
            snprintf(buf, sizeof(buf), " %s %.2f ms;", marks[k].first, ms);
            line += buf;
        }
        fprintf(stderr, "%s\n", line.c_str());
        for (auto& m : marks) cudaEventDestroy(m.second);
        marks.clear();
    }
};
thread_local SetupMarks g_marks;

int enqueue_setup(se3icp_ctx* c) {
    const RunConfig& cfg = c->cfg;
    const se3icp_params& p = c->params;
    cudaStream_t st = c->stream;
    int N = (int)c->n[0], M = (int)c->n[1];
    const double* rs = c->raw_view[0];
    const double* rt = c->raw_view[1];
    IterState* ds = c->dstate();

    g_marks.mark("begin", st);
    SE3_TRY(launch_init_state(ds, c->hist.as<unsigned int>(), st));
    c->launches += 1;
    if (cfg.has_se3) {
        if (cfg.with_cf) {  // .cpp:756-769: confidences from the raw depth, before normalisation
            SE3_TRY(launch_confidence(rs, N, c->conf[0].as<double>(), st));
            SE3_TRY(launch_confidence(rt, M, c->conf[1].as<double>(), st));
            c->launches += 2;
        }
        SE3_TRY(launch_sum_xyz(rs, N, c->psum[0].as<double>(), st));
        SE3_TRY(launch_sum_xyz(rt, M, c->psum[1].as<double>(), st));
        SE3_TRY(launch_maxdist(rs, N, c->psum[0].as<double>(), c->pmax[0].as<double>(), st));
        SE3_TRY(launch_maxdist(rt, M, c->psum[1].as<double>(), c->pmax[1].as<double>(), st));
        SE3_TRY(launch_normalise(rs, N, c->psum[0].as<double>(), c->pmax[0].as<double>(), c->pmax[1].as<double>(), N, M,
                                 p.scale_preprocessing, SE3ICP_SOURCE, ds, c->index[0].x.as<double>(),
                                 c->index[0].y.as<double>(), c->index[0].z.as<double>(), st));
        SE3_TRY(launch_normalise(rt, M, c->psum[1].as<double>(), c->pmax[0].as<double>(), c->pmax[1].as<double>(), N, M,
                                 p.scale_preprocessing, SE3ICP_TARGET, ds, c->index[1].x.as<double>(),
                                 c->index[1].y.as<double>(), c->index[1].z.as<double>(), st));
        c->launches += 6;
    } else {
        SE3_TRY(launch_aos_to_soa(rs, N, c->index[0].x.as<double>(), c->index[0].y.as<double>(),
                                  c->index[0].z.as<double>(), st));
        SE3_TRY(launch_aos_to_soa(rt, M, c->index[1].x.as<double>(), c->index[1].y.as<double>(),
                                  c->index[1].z.as<double>(), st));
        c->launches += 2;
    }
    g_marks.mark("normalise", st);
    SE3_TRY(c->index[1].build(st, &c->launches));
    const bool need_src_index = cfg.has_se3 || cfg.variant == SE3ICP_GICP;
    if (need_src_index) SE3_TRY(c->index[0].build(st, &c->launches));
    c->src_index_built = need_src_index;
    g_marks.mark("3-D indices", st);

    for (int w = 0; w < 2; w++) {
        FeatureArgs fa{};
        const bool shot = cfg.has_se3 && p.lrf_method == SE3ICP_LRF_SHOT;
        fa.k_lrf = cfg.has_se3 && !shot ? p.number_of_nn_for_LRF : 0;
        if (cfg.variant == SE3ICP_GICP)
            fa.k_nrm = p.knn_normals_gicp;  // .cpp:43 both clouds
        else if (cfg.variant == SE3ICP_PT2PL && w == 1)
            fa.k_nrm = p.knn_normals_pt2pl;  // .cpp:494,643 target only
        fa.want_cov = cfg.variant == SE3ICP_GICP;
        fa.gicp_eps = p.gicp_epsilon;
        fa.frame = c->frame[w].as<double>();
        fa.nrm = c->nrm[w].as<double>();
        fa.cov = c->cov[w].as<double>();
        fa.K = std::max(fa.k_lrf, fa.k_nrm);
        if (shot) {  // .cpp:593-594 (commented there): radius-support frames instead of the kNN ones; never reused
            c->feat[w].valid = false;
            SE3_TRY(launch_shot_lrf(c->index[w].view, p.lrf_radius, c->frame[w].as<double>(), nullptr, st));
            c->launches += 1;
        }
        fa.q_begin = 0;
        fa.q_end = 0x7fffffff;
        if (w == 0 && c->sharded) {  // source features are only needed for this rank's query range
            fa.q_begin = c->shard_begin;
            fa.q_end = c->shard_end;
        }
        // sharded pair: every rank needs ALL target features, but each computes only its slice of them; the slices
        // are exchanged below (one grouped NCCL broadcast per owner and plane, in place)
        const bool split_target = w == 1 && c->sharded && c->comm && c->comm_size > 1;
        auto slice = [&](int r, int& b, int& e) {
            int base = M / c->comm_size, rem = M % c->comm_size;
            b = r * base + (r < rem ? r : rem);
            e = b + base + (r < rem ? 1 : 0);
        };
        if (split_target) slice(c->comm_rank, fa.q_begin, fa.q_end);
        if (fa.K <= 0) continue;
        if (c->sharded) {
            SE3_TRY(c->knn_list.ensure(c->n[w] * sizeof(int)));
            SE3_TRY(c->knn_count.ensure(sizeof(int)));
            fa.active_list = c->knn_list.as<int>();
            fa.active_count = c->knn_count.as<int>();
        }
        se3icp_ctx::FeatureKey key;
        key.valid = !(w == 0 && c->sharded) && !shot;  // a sharded source only holds its own range; SHOT frames depend on the pair's scale
        key.n = c->n[w];
        key.k_lrf = fa.k_lrf, key.k_nrm = fa.k_nrm, key.want_cov = fa.want_cov, key.eps = fa.gicp_eps;
        const se3icp_ctx::FeatureKey& have = c->feat[w];
        // (never in a multi-rank run: every rank must take part in the exchange below)
        if (p.reuse_features && !split_target && have.valid && key.valid && have.n == key.n && have.k_lrf == key.k_lrf &&
            have.k_nrm == key.k_nrm && have.want_cov == key.want_cov && have.eps == key.eps) {
            c->feature_reuses += 1;  // computed by an earlier run on this very cloud (se3icp_swap_clouds)
            continue;
        }
        c->feat[w].valid = false;
        SE3_TRY(launch_knn_features(c->index[w].view, fa, st));
        c->launches += 1;
        g_marks.mark(w == 0 ? "kNN/features source" : "kNN/features target", st);
        if (split_target) {
            const NcclApi* nccl = nccl_api();
            if (!nccl) return SE3ICP_ERR_NCCL;
            ncclComm_t comm = (ncclComm_t)c->comm;
            struct Planes { double* base; int count; } sets[3] = {
                {fa.k_lrf > 0 ? fa.frame : nullptr, 9}, {fa.k_nrm > 0 ? fa.nrm : nullptr, 3},
                {fa.k_nrm > 0 && fa.want_cov ? fa.cov : nullptr, 6}};
            SE3_NCCL(nccl->GroupStart());
            for (int r = 0; r < c->comm_size; r++) {
                int b, e;
                slice(r, b, e);
                if (e <= b) continue;
                for (const Planes& ps : sets) {
                    if (!ps.base) continue;
                    for (int k = 0; k < ps.count; k++) {
                        double* ptr = ps.base + (size_t)k * (size_t)M + b;
                        SE3_NCCL(nccl->Broadcast(ptr, ptr, (size_t)(e - b), ncclDouble, r, comm, st));
                    }
                }
            }
            SE3_NCCL(nccl->GroupEnd());
            g_marks.mark("target feature exchange", st);
        }
        c->feat[w] = key;
    }
    // one gather record per target point for the reduction (after the features: it holds the normal / covariance)
    SE3_TRY(c->tgt_rec.ensure((size_t)M * kTargetRecordDoubles * sizeof(double)));
    SE3_TRY(launch_pack_target_records(c->index[1].view, cfg.variant == SE3ICP_PT2PL ? c->nrm[1].as<double>() : nullptr,
                                       cfg.variant == SE3ICP_GICP ? c->cov[1].as<double>() : nullptr, c->tgt_rec.as<double>(), st));
    c->tgt_rec_valid = true;
    c->launches += 1;
    if (cfg.has_se3)  // .cpp:597-626: weighting, 12 x M matrix and its search structure
        SE3_TRY(c->se3idx.build(c->index[1].view, c->frame[1].as<double>(), cfg.alpha, cfg.with_cf ? 1.0 : cfg.beta, ds, st,
                                &c->launches));
    SE3_CUDA(cudaMemsetAsync(c->corr_idx.ptr, 0xff, (size_t)N * sizeof(int), st));
    if (cfg.coherence || cfg.coherence_xyz) SE3_CUDA(cudaMemsetAsync(c->ref_d2nd.ptr, 0xff, (size_t)N * sizeof(double), st));  // NaN: not known
    if (cfg.trim_active && cfg.n_keep_target == 0) SE3_CUDA(cudaMemsetAsync(c->keep.ptr, 0, (size_t)N, st));
    SE3_CUDA(cudaMemsetAsync(c->tcount.ptr, 0, kTcountWords * sizeof(unsigned int), st));
    if (cfg.trim_active) SE3_CUDA(cudaMemsetAsync(c->thist.ptr, 0, (size_t)kTrimHistBins * sizeof(unsigned int), st));
    g_marks.mark("12-D index + resets", st);
    SE3_TRY(launch_mark_loop_start(ds, st));
    c->launches += 1;
    return 0;
}

int enqueue_iteration(se3icp_ctx* c, unsigned long long cond_handle = 0) {
    const RunConfig& cfg = c->cfg;
    cudaStream_t st = c->stream;
    SourceView S = c->source_view();
    TargetView T = c->target_view();
    CorrBuffers cb = c->corr_buffers(false);
    IterState* ds = c->dstate();
    SE3_TRY(launch_nn_filter(S, T, cfg, ds, cb, st));  // seeds the first pass of a run, then the coherence filter
    c->launches += 1;
    const int mode = c->params.nn_mode;
    if (cfg.has_se3 && (mode == SE3ICP_NN_BRUTE_F32 || mode == SE3ICP_NN_EXACT_F64)) {
        SE3_TRY(launch_nn_se3_brute(S, T, cfg, ds, cb, mode == SE3ICP_NN_EXACT_F64, st));
        SE3_TRY(launch_nn_se3_repair(S, T, cfg, ds, cb, st));
        c->launches += 2;
        if (!cfg.pure) {
            SE3_TRY(launch_nn_xyz(S, T, cfg, ds, cb, st));
            c->launches += 1;
        }
    } else {
        // AUTO / TREE: pruned traversals; one kernel serves both phases (the phase flag is device-side)
        SE3_TRY(launch_nn_search(S, T, cfg, ds, cb, st));
        c->launches += 1;
    }
    const bool multi = c->sharded && c->comm && c->comm_size > 1;
    const NcclApi* nccl = multi ? nccl_api() : nullptr;
    if (multi && !nccl) return SE3ICP_ERR_NCCL;
    ncclComm_t comm = (ncclComm_t)c->comm;
    if (cfg.trim_active && cfg.n_keep_target > 0) {
        if (cb.thist) {
            SE3_TRY(launch_trim_select(cfg, ds, cb, S.begin, S.end, st));
            c->launches += 2;
        } else {
            // global threshold: every radix pass all-reduces its 256-bin histogram; ties at the
            // threshold are granted in global index order (lower ranks first)
            const float* df = cb.distf + S.begin;
            int nl = S.end - S.begin;
            unsigned int* hist = c->hist.as<unsigned int>();
            for (int pass = 0; pass < 4; pass++) {
                SE3_TRY(launch_trim_hist(cfg, ds, df, nl, hist, pass, st));
                if (multi) SE3_NCCL(nccl->AllReduce(hist + pass * 256, hist + pass * 256, 256, ncclUint32, ncclSum, comm, st));
            }
            SE3_CUDA(cudaMemsetAsync(c->eq_total.ptr, 0, sizeof(int), st));
            SE3_TRY(launch_trim_count_eq(cfg, ds, df, nl, hist, c->block_eq.as<int>(), c->eq_total.as<int>(), st));
            const int* rank_eq = nullptr;
            if (multi) {
                SE3_NCCL(nccl->AllGather(c->eq_total.ptr, c->rank_eq.ptr, 1, ncclInt32, comm, st));
                rank_eq = c->rank_eq.as<int>();
            }
            SE3_TRY(launch_trim_apply(cfg, ds, df, nl, hist, c->block_eq.as<int>(), rank_eq, c->comm_rank, cb.keep + S.begin, st));
            c->launches += 6;
        }
    }
    // The last block of the reduction also solves, updates and decides (an iteration ends with this launch): always on
    // one GPU, and for a sharded pair whenever the ranks can all-reduce the record over peer memory inside that block.
    const bool p2p = multi && c->peer_ready && !(cfg.trim_active && cfg.n_keep_target > 0);  // (trimming: NCCL histograms)
    SolveFusion fuse{};
    fuse.enabled = !multi || p2p;
    fuse.peer.world = 1;
    if (p2p) {
        fuse.peer.mailboxes = c->mailbox_table.as<unsigned long long*>();
        fuse.peer.world = c->comm_size;
        fuse.peer.rank = c->comm_rank;
        fuse.peer.seq_base = c->peer_runs << 32;
    }
    fuse.history = c->history.as<double>();
    fuse.hist = cb.thist ? nullptr : c->hist.as<unsigned int>();
    fuse.cond_handle = cond_handle;
    SE3_TRY(launch_reduce(S, T, cfg, ds, cb, c->partials.as<double>(), fuse, st));
    c->launches += 1;
    if (multi && !p2p) {
        // NCCL path: one all-reduce of the 29-double record per iteration; every rank then runs the identical solve
        SE3_TRY(launch_sum_partials(c->partials.as<double>(), c->totals.as<double>(), st));
        SE3_NCCL(nccl->AllReduce(c->totals.ptr, c->totals.ptr, kReducePartials, ncclFloat64, ncclSum, comm, st));
        SE3_TRY(launch_solve_update(cfg, ds, c->totals.as<double>(), 1, c->history.as<double>(), c->hist.as<unsigned int>(),
                                    cond_handle, st));
        c->launches += 2;
    }
    return 0;
}

// The loop graph is the default; SE3ICP_USE_GRAPH=0/1 overrides the parameter.  Kernels inside a graph that
// contains a conditional node cannot be profiled by Nsight Compute ("not supported for profiling"), so
// when the process runs under ncu the host-driven loop is used and every kernel stays visible.
bool want_graph(const se3icp_params* p) {
    static const int forced = [] {
        const char* e = getenv("SE3ICP_USE_GRAPH");
        if (e && *e) return atoi(e) != 0 ? 1 : 0;
        if (getenv("NV_COMPUTE_PROFILER_PERFWORKS_DIR") || getenv("NV_NSIGHT_INJECTION_PORT_BASE")) {
            fprintf(stderr, "[se3icp] profiler detected: using the host-driven iteration loop (graph kernels are not profilable)\n");
            return 0;
        }
        return -1;
    }();
    return forced >= 0 ? forced != 0 : p->use_graph != 0;
}

void release_loop_graph(se3icp_ctx* c) {
    if (c->loop_exec) cudaGraphExecDestroy(c->loop_exec);
    if (c->loop_graph) cudaGraphDestroy(c->loop_graph);
    c->loop_exec = nullptr;
    c->loop_graph = nullptr;
}

// The loop graph of a run: a conditional WHILE node whose body is the captured iteration.  Capturing costs ~10 us;
// instantiating ~100 us (measured, profiles/experiments/cond_update_test.cu), so the executable graph is kept by the
// context and only re-parameterised (cudaGraphExecUpdate, ~3 us) from the freshly captured one — pointers, sizes and
// grids differ from pair to pair, the topology does not.  It is re-instantiated when the update is refused (another
// launch sequence: trimming switched on or off, another search mode).
int build_loop_graph(se3icp_ctx* c) {
    cudaGraph_t graph = nullptr;
    SE3_CUDA(cudaGraphCreate(&graph, 0));
    cudaGraphConditionalHandle handle;
    cudaGraphNodeParams np = {};
    cudaGraphNode_t node;
    cudaError_t e = cudaGraphConditionalHandleCreate(&handle, graph, 1, cudaGraphCondAssignDefault);
    if (e == cudaSuccess) {
        np.type = cudaGraphNodeTypeConditional;
        np.conditional.handle = handle;
        np.conditional.type = cudaGraphCondTypeWhile;
        np.conditional.size = 1;
        e = cudaGraphAddNode(&node, graph, nullptr, 0, &np);
    }
    if (e == cudaSuccess)
        e = cudaStreamBeginCaptureToGraph(c->stream, np.conditional.phGraph_out[0], nullptr, nullptr, 0,
                                          cudaStreamCaptureModeThreadLocal);
    if (e != cudaSuccess) {
        cudaGraphDestroy(graph);
        SE3_CUDA(e);
    }
    long long before = c->launches;
    int rc = enqueue_iteration(c, (unsigned long long)handle);  // the last block of reduce_kernel sets the loop condition
    cudaGraph_t captured = nullptr;
    e = cudaStreamEndCapture(c->stream, &captured);
    c->launches_per_iter = c->launches - before;
    c->launches = before;
    if (rc || e != cudaSuccess) {
        cudaGraphDestroy(graph);
        if (rc) return rc;
        SE3_CUDA(e);
    }
    if (c->loop_exec) {
        cudaGraphExecUpdateResultInfo info;
        if (cudaGraphExecUpdate(c->loop_exec, graph, &info) != cudaSuccess) {
            cudaGetLastError();  // not an error of the run: fall back to a fresh executable
            cudaGraphExecDestroy(c->loop_exec);
            c->loop_exec = nullptr;
        } else {
            c->graph_updates += 1;
        }
    }
    if (!c->loop_exec) {
        e = cudaGraphInstantiate(&c->loop_exec, graph, 0);
        if (e != cudaSuccess) {
            cudaGraphDestroy(graph);
            SE3_CUDA(e);
        }
        c->graph_instantiations += 1;
    }
    if (c->loop_graph) cudaGraphDestroy(c->loop_graph);
    c->loop_graph = graph;
    return 0;
}

int check_ctx(se3icp_ctx* c) {
    if (!c) {
        set_last_error("null context");
        return SE3ICP_ERR_ARG;
    }
    SE3_CUDA(cudaSetDevice(c->device));
    return 0;
}

}  // namespace

// ================================================================================================
extern "C" {

int se3icp_abi_version(void) { return SE3ICP_ABI_VERSION; }
const char* se3icp_last_error(void) { return se3::g_err; }

void se3icp_default_params(se3icp_params* p) {  // reference ctor .cpp:334-348
    if (!p) return;
    memset(p, 0, sizeof(*p));
    p->variant = SE3ICP_PT2PL;
    p->entry = SE3ICP_RUN_SE3_ICP;
    p->max_num_iterations = 150;
    p->max_num_se3_iterations = 20;
    p->number_of_nn_for_LRF = 30;
    p->knn_normals_pt2pl = 30;
    p->knn_normals_gicp = 20;
    p->trim_keep_largest = 1;  /* PCL's comparator (pc1.distance > pc2.distance): see se3icp.h */
    p->mse = 0.00001;
    p->mse_switch_error = 0.001;
    p->estimated_overlap = 1.0;
    p->alpha_rot = 3.0;
    p->beta_transl = 1.0;
    p->scale_preprocessing = 3.0;
    p->gicp_epsilon = 1e-3;
    p->nn_mode = SE3ICP_NN_AUTO;
    p->use_graph = 1;  /* whole loop as one CUDA graph (conditional WHILE node) */
    p->record_history = 0;
    p->nn_coherence = 1;
    p->reuse_features = 1;
    p->lrf_method = SE3ICP_LRF_TOLDI;
    p->reserved0 = 0;
    p->lrf_radius = 0.8;
}

int se3icp_create(int device, void* stream, se3icp_ctx** out) {
    if (!out) return SE3ICP_ERR_ARG;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0 || device < 0 || device >= count) {
        set_last_error("no usable CUDA device %d (count %d, %s)", device, count, cudaGetErrorString(e));
        return SE3ICP_ERR_NO_DEVICE;
    }
    cudaDeviceProp prop;
    SE3_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_last_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
        return SE3ICP_ERR_NO_DEVICE;
    }
    SE3_CUDA(cudaSetDevice(device));
    se3icp_ctx* c = new se3icp_ctx();
    c->device = device;
    if (stream) {
        c->stream = (cudaStream_t)stream;
    } else {
        if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete c;
            set_last_error("cudaStreamCreate failed");
            return SE3ICP_ERR_CUDA;
        }
        c->own_stream = true;
    }
    if (cudaMallocHost((void**)&c->h_state, sizeof(IterState)) != cudaSuccess ||
        cudaMallocHost((void**)&c->h_flag, 64) != cudaSuccess || cudaEventCreate(&c->ev_begin) != cudaSuccess ||
        cudaEventCreate(&c->ev_setup) != cudaSuccess ||
        // blocking-sync: a thread waiting for a run sleeps instead of spinning (8 ranks x 2 enqueue threads share the
        // host cores of one box; spinning waiters cost 14 % of the 8-GPU batch throughput, measured)
        cudaEventCreateWithFlags(&c->ev_end, cudaEventBlockingSync) != cudaSuccess) {
        set_last_error("context allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        se3icp_destroy(c);
        return SE3ICP_ERR_CUDA;
    }
    *out = c;
    return SE3ICP_OK;
}

int se3icp_destroy(se3icp_ctx* c) {
    if (!c) return SE3ICP_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    if (c->comm && c->comm_owned) se3icp_comm_destroy(c);
    peer_teardown(c);
    release_loop_graph(c);
    if (c->h_state) cudaFreeHost(c->h_state);
    if (c->h_flag) cudaFreeHost(c->h_flag);
    if (c->ev_begin) cudaEventDestroy(c->ev_begin);
    if (c->ev_setup) cudaEventDestroy(c->ev_setup);
    if (c->ev_end) cudaEventDestroy(c->ev_end);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
    return SE3ICP_OK;
}

int se3icp_synchronize(se3icp_ctx* c) {
    SE3_TRY(check_ctx(c));
    SE3_CUDA(cudaStreamSynchronize(c->stream));
    return SE3ICP_OK;
}

// entry points that would reallocate or overwrite what an enqueued run still uses
#define SE3_NOT_PENDING(name)                                                     \
    do {                                                                          \
        if (c->run_pending) {                                                     \
            set_last_error(name ": a run is pending (call se3icp_run_finish)");   \
            return SE3ICP_ERR_STATE;                                              \
        }                                                                         \
    } while (0)

int se3icp_set_cloud(se3icp_ctx* c, int which, const double* xyz, size_t n, int append) {
    SE3_TRY(check_ctx(c));
    SE3_NOT_PENDING("se3icp_set_cloud");
    if ((which != SE3ICP_SOURCE && which != SE3ICP_TARGET) || (!xyz && n > 0)) {
        set_last_error("se3icp_set_cloud: bad argument");
        return SE3ICP_ERR_ARG;
    }
    if (n > 0x7fffffffULL / 4) return SE3ICP_ERR_UNSUPPORTED;
    size_t old = append ? c->n[which] : 0;
    if (append && c->raw_view[which] != c->raw[which].ptr && old > 0) {
        set_last_error("cannot append to a caller-owned device cloud");
        return SE3ICP_ERR_STATE;
    }
    size_t total = old + n;
    SE3_TRY(c->raw[which].ensure_keep(total * 3 * sizeof(double), old * 3 * sizeof(double), c->stream));
    if (n > 0)
        SE3_CUDA(cudaMemcpyAsync(c->raw[which].as<double>() + old * 3, xyz, n * 3 * sizeof(double), cudaMemcpyHostToDevice,
                                 c->stream));
    c->raw_view[which] = c->raw[which].as<double>();
    c->n[which] = total;
    c->feat[which].valid = false;
    return SE3ICP_OK;
}

int se3icp_set_cloud_device(se3icp_ctx* c, int which, const double* d_xyz, size_t n) {
    SE3_TRY(check_ctx(c));
    SE3_NOT_PENDING("se3icp_set_cloud_device");
    if ((which != SE3ICP_SOURCE && which != SE3ICP_TARGET) || !d_xyz || n == 0) return SE3ICP_ERR_ARG;
    c->raw_view[which] = d_xyz;
    c->n[which] = n;
    c->feat[which].valid = false;
    return SE3ICP_OK;
}

int se3icp_swap_clouds(se3icp_ctx* c) {
    SE3_TRY(check_ctx(c));
    SE3_NOT_PENDING("se3icp_swap_clouds");
    c->raw[0].swap(c->raw[1]);
    std::swap(c->raw_view[0], c->raw_view[1]);
    std::swap(c->n[0], c->n[1]);
    c->index[0].swap(c->index[1]);
    c->frame[0].swap(c->frame[1]);
    c->nrm[0].swap(c->nrm[1]);
    c->cov[0].swap(c->cov[1]);
    c->conf[0].swap(c->conf[1]);
    c->psum[0].swap(c->psum[1]);
    c->pmax[0].swap(c->pmax[1]);
    std::swap(c->feat[0], c->feat[1]);
    c->src_index_built = false;
    return SE3ICP_OK;
}

static int run_async_impl(se3icp_ctx* c, const se3icp_params* p);

int se3icp_run_async(se3icp_ctx* c, const se3icp_params* p) {
    SE3_TRY(check_ctx(c));
    c->sharded = false;
    return run_async_impl(c, p);
}

static int run_async_impl(se3icp_ctx* c, const se3icp_params* p) {
    if (!p) return SE3ICP_ERR_ARG;
    SE3_NOT_PENDING("se3icp_run_async");
    if (c->n[0] == 0 || c->n[1] == 0 || !c->raw_view[0] || !c->raw_view[1]) {
        set_last_error("source/target cloud not set");
        return SE3ICP_ERR_STATE;
    }
    if (p->number_of_nn_for_LRF > SE3ICP_MAX_KNN || p->knn_normals_gicp > SE3ICP_MAX_KNN ||
        p->knn_normals_pt2pl > SE3ICP_MAX_KNN) {
        set_last_error("kNN sizes above %d are not supported", SE3ICP_MAX_KNN);
        return SE3ICP_ERR_UNSUPPORTED;
    }
    if (p->lrf_method != SE3ICP_LRF_TOLDI && p->lrf_method != SE3ICP_LRF_SHOT) {
        set_last_error("bad lrf_method %d", p->lrf_method);
        return SE3ICP_ERR_ARG;
    }
    if (p->lrf_method == SE3ICP_LRF_SHOT && p->entry != SE3ICP_RUN_ICP) {
        if (!(p->lrf_radius > 0.0)) {
            set_last_error("SHOT frame: lrf_radius must be positive");
            return SE3ICP_ERR_ARG;
        }
        if (c->sharded) {
            set_last_error("the sharded pair computes TOLDI frames only");
            return SE3ICP_ERR_UNSUPPORTED;
        }
    }
    if (p->nn_mode < SE3ICP_NN_AUTO || p->nn_mode > SE3ICP_NN_TREE) {
        set_last_error("bad nn_mode %d", p->nn_mode);
        return SE3ICP_ERR_ARG;
    }
    SE3_TRY(fill_config(c, p));
    SE3_TRY(alloc_run(c));
    c->launches = 0;
    c->feature_reuses = 0;
    SE3_CUDA(cudaEventRecord(c->ev_begin, c->stream));
    SE3_TRY(enqueue_setup(c));
    SE3_CUDA(cudaEventRecord(c->ev_setup, c->stream));
    // Iterations: the stop/phase decision lives on the device (tail of reduce_kernel).
    c->graph_run = false;
    const bool multi_rank = c->sharded && c->comm && c->comm_size > 1;
    const bool p2p = multi_rank && c->peer_ready && !(c->cfg.trim_active && c->cfg.n_keep_target > 0);
    if (multi_rank) c->peer_runs += 1;  // collective call: the same count on every rank -> unique sequence words
    if (want_graph(p) && (!multi_rank || p2p)) {
        // The whole loop is ONE graph launch: a conditional WHILE node whose body is the captured iteration;
        // its last kernel sets the condition from the device-side done flag.  The host never round-trips.
        SE3_TRY(build_loop_graph(c));
        SE3_CUDA(cudaGraphLaunch(c->loop_exec, c->stream));
        c->graph_run = true;
    } else {
        // host-driven: poll the 4-byte done flag once per iteration.  Every kernel early-outs once done
        // is set, so over-issuing is harmless.
        const long hard_cap = 1000000;
        for (long it = 0; it < hard_cap; ++it) {
            SE3_TRY(enqueue_iteration(c));
            SE3_CUDA(cudaMemcpyAsync(c->h_flag, &c->dstate()->done, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
            SE3_CUDA(cudaStreamSynchronize(c->stream));
            if (*c->h_flag) break;
        }
    }
    SE3_TRY(launch_finalize(c->cfg, c->dstate(), c->stream));
    c->launches += 1;
    SE3_CUDA(cudaMemcpyAsync(c->h_state, c->dstate(), sizeof(IterState), cudaMemcpyDeviceToHost, c->stream));
    SE3_CUDA(cudaEventRecord(c->ev_end, c->stream));
    c->run_pending = true;
    return SE3ICP_OK;
}

int se3icp_run_finish(se3icp_ctx* c, double* T_out, se3icp_stats* stats) {
    SE3_TRY(check_ctx(c));
    if (!c->run_pending) {
        set_last_error("no run pending");
        return SE3ICP_ERR_STATE;
    }
    c->run_pending = false;  // whatever happens below, the context accepts calls again
    static const bool spin = [] {
        const char* e = getenv("SE3ICP_SPIN_WAIT");
        return e && atoi(e) != 0;
    }();
    if (spin)
        SE3_CUDA(cudaStreamSynchronize(c->stream));
    else
        SE3_CUDA(cudaEventSynchronize(c->ev_end));  // last thing run_async enqueued (after the copy of the state)
    g_marks.report(c->comm_rank);
    if (c->h_state->peer_timeout) {
        set_last_error("sharded pair: the record of a peer rank did not arrive within 20 s (rank %d of %d)", c->comm_rank,
                       c->comm_size);
        return SE3ICP_ERR_NCCL;
    }
    const IterState& hs = *c->h_state;
    if (T_out) memcpy(T_out, hs.T_final, 16 * sizeof(double));
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->num_iterations = hs.iter;
        stats->num_pure_se3_iterations = c->cfg.has_se3 ? hs.se3_iters : -1;
        stats->scaling_factor = hs.scale;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, c->ev_begin, c->ev_end);
        stats->time_total_ms = ms;
        cudaEventElapsedTime(&ms, c->ev_begin, c->ev_setup);
        stats->time_setup_ms = ms;
        stats->time_se3_correspondence_search_ms = (double)hs.t_corr_ns / 1e6;
        stats->time_se3_phase_search_ms = (double)hs.t_corr_se3_ns / 1e6;
        stats->time_before_pure_icp_ms = stats->time_total_ms;  // .cpp:957-958 measures the whole call
        stats->exact_repairs = hs.total_repairs;
        stats->kernel_launches = c->launches + (c->graph_run ? c->launches_per_iter * (long long)hs.iter : 0);
        stats->feature_reuses = c->feature_reuses;
        stats->queries_searched = (long long)hs.searched_total;
        stats->graph_instantiations = c->graph_instantiations;
        stats->loop_was_graph = c->graph_run ? 1 : 0;
    }
    return SE3ICP_OK;
}

int se3icp_run(se3icp_ctx* c, const se3icp_params* p, double* T_out, se3icp_stats* stats) {
    SE3_TRY(se3icp_run_async(c, p));
    return se3icp_run_finish(c, T_out, stats);
}

int se3icp_get_history(se3icp_ctx* c, double* T_hist, int max_entries, int* n_out) {
    SE3_TRY(check_ctx(c));
    if (!n_out) return SE3ICP_ERR_ARG;
    int cnt = c->h_state ? c->h_state->hist_count : 0;
    if (!c->cfg.record_history) cnt = 0;
    *n_out = cnt;
    int k = std::min(cnt, max_entries);
    if (k > 0 && T_hist) {
        SE3_CUDA(cudaMemcpyAsync(T_hist, c->history.ptr, (size_t)k * 16 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        SE3_CUDA(cudaStreamSynchronize(c->stream));
    }
    return SE3ICP_OK;
}

int se3icp_get_correspondences(se3icp_ctx* c, int32_t* tgt_idx, double* dist, size_t n) {
    SE3_TRY(check_ctx(c));
    if (n > c->n[0] || !c->corr_idx.ptr) return SE3ICP_ERR_ARG;
    if (tgt_idx) SE3_CUDA(cudaMemcpyAsync(tgt_idx, c->corr_idx.ptr, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (dist) SE3_CUDA(cudaMemcpyAsync(dist, c->corr_dist.ptr, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    SE3_CUDA(cudaStreamSynchronize(c->stream));
    return SE3ICP_OK;
}

// frames16: [alpha R | beta p] of the normalised cloud, as the reference leaves them after set-up
// (source: additionally left-multiplied by the accumulated estimate, .cpp:713-716)
int se3icp_get_se3_cloud(se3icp_ctx* c, int which, double* frames16, size_t n) {
    SE3_TRY(check_ctx(c));
    if ((which != 0 && which != 1) || !frames16 || n > c->n[which] || !c->cfg.has_se3 || !c->frame[which].ptr)
        return SE3ICP_ERR_ARG;
    size_t nn = c->n[which];
    std::vector<double> fr(9 * nn), x(nn), y(nn), z(nn);
    SE3_CUDA(cudaMemcpyAsync(fr.data(), c->frame[which].ptr, 9 * nn * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    SE3_CUDA(cudaMemcpyAsync(x.data(), c->index[which].x.ptr, nn * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    SE3_CUDA(cudaMemcpyAsync(y.data(), c->index[which].y.ptr, nn * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    SE3_CUDA(cudaMemcpyAsync(z.data(), c->index[which].z.ptr, nn * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    SE3_CUDA(cudaStreamSynchronize(c->stream));
    const double a = c->cfg.alpha, b = c->cfg.beta;
    const double* Tt = c->h_state->T_total;
    for (size_t i = 0; i < n; i++) {
        double X[16] = {0};
        for (int col = 0; col < 3; col++)
            for (int r = 0; r < 3; r++) X[4 * r + col] = a * fr[(size_t)(3 * col + r) * nn + i];
        X[3] = b * x[i], X[7] = b * y[i], X[11] = b * z[i];
        X[15] = 1.0;
        double* o = frames16 + 16 * i;
        if (which == SE3ICP_SOURCE) {
            for (int r = 0; r < 4; r++)
                for (int col = 0; col < 4; col++) {
                    double s = 0;
                    for (int k = 0; k < 4; k++) s += Tt[4 * r + k] * X[4 * k + col];
                    o[4 * r + col] = s;
                }
        } else {
            memcpy(o, X, sizeof(X));
        }
    }
    return SE3ICP_OK;
}

// ---- batch of independent pairs: the contexts run concurrently on their streams, fed by at most two host threads.
// A run is one asynchronous enqueue (set-up kernels + one graph launch), so a thread walks round-robin over its
// contexts: collect the result of the pair a context finished, hand it the next one.  (Round 1 used one polling
// thread per context: 16 threads per GPU x 8 ranks fought over the host cores and cost 8 % at 8 GPUs.)
static int run_batch_impl(se3icp_ctx** ctxs, int n_ctx, int n_pairs, const double* const* src, const size_t* n_src,
                          const double* const* tgt, const size_t* n_tgt, const se3icp_params* p, double* T_out,
                          se3icp_stats* stats, bool device_inputs) {
    if (!ctxs || n_ctx <= 0 || n_pairs < 0 || !src || !tgt || !n_src || !n_tgt || !p || !T_out) return SE3ICP_ERR_ARG;
    int n_threads = n_ctx < 2 ? n_ctx : 2;
    if (const char* e = getenv("SE3ICP_BATCH_THREADS")) n_threads = std::max(1, std::min(n_ctx, atoi(e)));
    std::vector<int> rc(n_threads, 0);
    std::vector<std::string> errs(n_threads);
    auto worker = [&](int t) {
        // pair pi runs on context pi % n_ctx; thread t owns the contexts t, t + n_threads, ...
        auto finish = [&](int pi) { return se3icp_run_finish(ctxs[pi % n_ctx], T_out + 16 * (size_t)pi, stats ? stats + pi : nullptr); };
        int r = 0;
        const int rounds = (n_pairs + n_ctx - 1) / n_ctx;
        for (int round = 0; round <= rounds && !r; round++) {
            for (int ci = t; ci < n_ctx && !r; ci += n_threads) {
                se3icp_ctx* c = ctxs[ci];
                const int prev = (round - 1) * n_ctx + ci, pi = round * n_ctx + ci;
                if (round > 0 && prev < n_pairs) r = finish(prev);
                if (r || pi >= n_pairs) continue;
                if (device_inputs) {
                    r = se3icp_set_cloud_device(c, SE3ICP_SOURCE, src[pi], n_src[pi]);
                    if (!r) r = se3icp_set_cloud_device(c, SE3ICP_TARGET, tgt[pi], n_tgt[pi]);
                } else {
                    r = se3icp_set_cloud(c, SE3ICP_SOURCE, src[pi], n_src[pi], 0);
                    if (!r) r = se3icp_set_cloud(c, SE3ICP_TARGET, tgt[pi], n_tgt[pi], 0);
                }
                if (!r) r = se3icp_run_async(c, p);
            }
        }
        if (r) {
            rc[t] = r;
            errs[t] = se3icp_last_error();
            for (int ci = t; ci < n_ctx; ci += n_threads)  // leave no run pending behind an error
                if (ctxs[ci]->run_pending) se3icp_run_finish(ctxs[ci], nullptr, nullptr);
        }
    };
    if (n_threads == 1) {
        worker(0);
    } else {
        std::vector<std::thread> th;
        for (int t = 0; t < n_threads; t++) th.emplace_back(worker, t);
        for (auto& t : th) t.join();
    }
    for (int t = 0; t < n_threads; t++)
        if (rc[t]) {
            set_last_error("%s", errs[t].c_str());
            return rc[t];
        }
    return SE3ICP_OK;
}

int se3icp_run_sequence(se3icp_ctx* c, const double* const* scans, const size_t* n_points, int n_scans,
                        const se3icp_params* p, int device_inputs, double* T_out, se3icp_stats* stats) {
    SE3_TRY(check_ctx(c));
    if (!scans || !n_points || n_scans < 2 || !p || !T_out) return SE3ICP_ERR_ARG;
    auto load = [&](int which, int i) {
        return device_inputs ? se3icp_set_cloud_device(c, which, scans[i], n_points[i])
                             : se3icp_set_cloud(c, which, scans[i], n_points[i], 0);
    };
    SE3_TRY(load(SE3ICP_SOURCE, 0));
    for (int i = 0; i + 1 < n_scans; i++) {
        SE3_TRY(se3icp_swap_clouds(c));  // scan i, source of the previous pair (features kept), becomes the target
        SE3_TRY(load(SE3ICP_SOURCE, i + 1));
        SE3_TRY(se3icp_run(c, p, T_out + 16 * (size_t)i, stats ? stats + i : nullptr));
    }
    return SE3ICP_OK;
}

int se3icp_run_batch(se3icp_ctx** ctxs, int n_ctx, int n_pairs, const double* const* src, const size_t* n_src,
                     const double* const* tgt, const size_t* n_tgt, const se3icp_params* p, double* T_out,
                     se3icp_stats* stats) {
    return run_batch_impl(ctxs, n_ctx, n_pairs, src, n_src, tgt, n_tgt, p, T_out, stats, false);
}

int se3icp_run_batch_device(se3icp_ctx** ctxs, int n_ctx, int n_pairs, const double* const* d_src, const size_t* n_src,
                            const double* const* d_tgt, const size_t* n_tgt, const se3icp_params* p, double* T_out,
                            se3icp_stats* stats) {
    return run_batch_impl(ctxs, n_ctx, n_pairs, d_src, n_src, d_tgt, n_tgt, p, T_out, stats, true);
}

int se3icp_comm_unique_id(void* id) {
    if (!id) return SE3ICP_ERR_ARG;
    const NcclApi* nccl = nccl_api();
    if (!nccl) return SE3ICP_ERR_NCCL;
    ncclUniqueId uid;
    SE3_NCCL(nccl->GetUniqueId(&uid));
    static_assert(sizeof(ncclUniqueId) == SE3ICP_COMM_ID_BYTES, "ncclUniqueId size");
    memcpy(id, &uid, sizeof(uid));
    return SE3ICP_OK;
}

int se3icp_comm_init(se3icp_ctx* c, int n_ranks, int rank, const void* id) {
    SE3_TRY(check_ctx(c));
    if (!id || n_ranks < 1 || rank < 0 || rank >= n_ranks) return SE3ICP_ERR_ARG;
    const NcclApi* nccl = nccl_api();
    if (!nccl) return SE3ICP_ERR_NCCL;
    if (c->comm && c->comm_owned) nccl->CommDestroy((ncclComm_t)c->comm);
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    ncclComm_t comm;
    SE3_NCCL(nccl->CommInitRank(&comm, n_ranks, uid, rank));
    c->comm = comm;
    c->comm_owned = true;
    c->comm_rank = rank;
    c->comm_size = n_ranks;
    SE3_TRY(peer_setup(c));  // collective, like the communicator itself
    return SE3ICP_OK;
}

int se3icp_comm_destroy(se3icp_ctx* c) {
    SE3_TRY(check_ctx(c));
    cudaStreamSynchronize(c->stream);
    peer_teardown(c);
    if (c->comm && c->comm_owned) {
        const NcclApi* nccl = nccl_api();
        if (nccl) nccl->CommDestroy((ncclComm_t)c->comm);
    }
    c->comm = nullptr;
    c->comm_owned = false;
    c->comm_rank = 0;
    c->comm_size = 1;
    return SE3ICP_OK;
}

int se3icp_comm_info(se3icp_ctx* c, int* rank_out, int* n_ranks_out) {
    SE3_TRY(check_ctx(c));
    if (rank_out) *rank_out = c->comm ? c->comm_rank : 0;
    if (n_ranks_out) *n_ranks_out = c->comm ? c->comm_size : 1;
    return SE3ICP_OK;
}

int se3icp_run_sharded(se3icp_ctx* c, const se3icp_params* p, size_t src_begin, size_t src_end, void* nccl_comm,
                       int rank, int n_ranks, double* T_out, se3icp_stats* stats) {
    SE3_TRY(check_ctx(c));
    if (src_begin > src_end || src_end > c->n[0]) {
        set_last_error("bad source range [%zu, %zu) of %zu", src_begin, src_end, c->n[0]);
        return SE3ICP_ERR_ARG;
    }
    if (nccl_comm && nccl_comm != c->comm) {  // caller-owned communicator, first use
        if (c->comm) se3icp_comm_destroy(c);
        c->comm = nccl_comm;
        c->comm_owned = false;
        c->comm_rank = rank;
        c->comm_size = n_ranks;
        SE3_TRY(peer_setup(c));
    }
    if (!c->comm && (src_begin != 0 || src_end != c->n[0])) {
        set_last_error("a partial source range needs a communicator (se3icp_comm_init)");
        return SE3ICP_ERR_STATE;
    }
    c->sharded = true;
    c->shard_begin = (int)src_begin;
    c->shard_end = (int)src_end;
    int rc = run_async_impl(c, p);
    if (rc == SE3ICP_OK) rc = se3icp_run_finish(c, T_out, stats);
    c->sharded = false;
    return rc;
}

int se3icp_time_stage(se3icp_ctx* c, int stage, int repeats, double* ms_avg) {
    SE3_TRY(check_ctx(c));
    if (!ms_avg || repeats <= 0 || stage < SE3ICP_STAGE_NN_SE3 || stage > SE3ICP_STAGE_KNN_TARGET) return SE3ICP_ERR_ARG;
    if (c->run_pending || c->n[0] == 0 || c->n[1] == 0 || !c->state.ptr) {
        set_last_error("se3icp_time_stage needs a finished se3icp_run on this context");
        return SE3ICP_ERR_STATE;
    }
    if ((stage == SE3ICP_STAGE_NN_SE3 && !c->cfg.has_se3) || (stage == SE3ICP_STAGE_NN_XYZ && c->cfg.pure))
        return SE3ICP_ERR_STATE;
    cudaStream_t st = c->stream;
    IterState saved = *c->h_state;
    IterState tmp = saved;
    tmp.done = 0;
    tmp.repair_count = 0;
    tmp.switch_icp = stage == SE3ICP_STAGE_NN_SE3 ? 0 : 1;
    tmp.T_change = 1e7;  // time the full search, not the coherence shortcut
    SE3_CUDA(cudaMemcpyAsync(c->dstate(), &tmp, sizeof(IterState), cudaMemcpyHostToDevice, st));
    SourceView S = c->source_view();
    TargetView T = c->target_view();
    CorrBuffers cb = c->corr_buffers(false);
    RunConfig cfg = c->cfg;
    cfg.pure = 0;
    cfg.coherence = 0;      // time the full search over every query, not the coherence shortcut
    cfg.coherence_xyz = 0;
    cudaEvent_t e0, e1;
    SE3_CUDA(cudaEventCreate(&e0));
    SE3_CUDA(cudaEventCreate(&e1));
    double total = 0.0;
    for (int r = -1; r < repeats; r++) {  // r == -1 is a warm-up launch
        if (stage == SE3ICP_STAGE_NN_SE3)
            SE3_CUDA(cudaMemsetAsync(&c->dstate()->repair_count, 0, sizeof(int), st));
        if (stage == SE3ICP_STAGE_NN_SE3 || stage == SE3ICP_STAGE_NN_XYZ)  // cold search: no remembered matches
            SE3_CUDA(cudaMemsetAsync(c->corr_idx.ptr, 0xff, c->n[0] * sizeof(int), st));
        SE3_CUDA(cudaEventRecord(e0, st));
        switch (stage) {
            case SE3ICP_STAGE_NN_SE3:
                if (c->params.nn_mode == SE3ICP_NN_BRUTE_F32 || c->params.nn_mode == SE3ICP_NN_EXACT_F64) {
                    SE3_TRY(launch_nn_se3_brute(S, T, cfg, c->dstate(), cb, 0, st));
                } else {  // cold pass = seeding (nn_filter_kernel) + search
                    SE3_TRY(launch_nn_filter(S, T, cfg, c->dstate(), cb, st));
                    SE3_TRY(launch_nn_se3_tree(S, T, cfg, c->dstate(), cb, st));
                }
                break;
            case SE3ICP_STAGE_NN_XYZ:
                SE3_TRY(launch_nn_filter(S, T, cfg, c->dstate(), cb, st));
                SE3_TRY(launch_nn_xyz(S, T, cfg, c->dstate(), cb, st));
                break;
            case SE3ICP_STAGE_REDUCE:
                SE3_TRY(launch_reduce(S, T, cfg, c->dstate(), cb, c->partials.as<double>(), SolveFusion{}, st));
                break;
            default: {
                FeatureArgs fa{};
                fa.k_lrf = cfg.has_se3 ? c->params.number_of_nn_for_LRF : 0;
                fa.k_nrm = cfg.variant == SE3ICP_GICP ? c->params.knn_normals_gicp
                                                      : (cfg.variant == SE3ICP_PT2PL ? c->params.knn_normals_pt2pl : 0);
                fa.want_cov = cfg.variant == SE3ICP_GICP;
                fa.gicp_eps = c->params.gicp_epsilon;
                fa.frame = c->frame[1].as<double>();
                fa.nrm = c->nrm[1].as<double>();
                fa.cov = c->cov[1].as<double>();
                fa.K = std::max(fa.k_lrf, fa.k_nrm);
                fa.q_begin = 0;
                fa.q_end = 0x7fffffff;
                if (fa.K <= 0) return SE3ICP_ERR_STATE;
                SE3_TRY(launch_knn_features(c->index[1].view, fa, st));
            }
        }
        SE3_CUDA(cudaEventRecord(e1, st));
        SE3_CUDA(cudaEventSynchronize(e1));
        float ms = 0.f;
        SE3_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (r >= 0) total += ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    SE3_CUDA(cudaMemcpyAsync(c->dstate(), &saved, sizeof(IterState), cudaMemcpyHostToDevice, st));
    SE3_CUDA(cudaStreamSynchronize(st));
    *ms_avg = total / repeats;
    return SE3ICP_OK;
}

// ================================================================================================
// stage-level entry points
// ================================================================================================
namespace {

// The stage-level entry points borrow the context's cloud slots, feature planes and iteration state as scratch.
// Whatever clouds the context held are gone afterwards: the scope guard marks them unset, so a later se3icp_run
// without a fresh se3icp_set_cloud fails with SE3ICP_ERR_STATE instead of reading half-overwritten buffers.
struct StageScope {
    se3icp_ctx* c;
    explicit StageScope(se3icp_ctx* ctx) : c(ctx) { c->tgt_rec_valid = false; }  // stage data lives in the planes
    ~StageScope() {
        for (int w = 0; w < 2; w++) {
            c->n[w] = 0;
            c->raw_view[w] = nullptr;
            c->feat[w].valid = false;
        }
        c->src_index_built = false;
    }
};

int upload_cloud_and_index(se3icp_ctx* c, int w, const double* xyz, size_t n) {
    SE3_TRY(se3icp_set_cloud(c, w, xyz, n, 0));
    SE3_TRY(c->index[w].reserve((int)n));
    SE3_TRY(launch_aos_to_soa(c->raw_view[w], (int)n, c->index[w].x.as<double>(), c->index[w].y.as<double>(),
                              c->index[w].z.as<double>(), c->stream));
    SE3_TRY(c->index[w].build(c->stream, nullptr));
    return 0;
}

// host row-major [n][d] -> device planes [d][n]
int upload_planes(se3icp_ctx* c, DeviceBuf& dst, const double* rows, size_t n, int d, int stride, int col0) {
    std::vector<double> planes((size_t)d * n);
    for (size_t i = 0; i < n; i++)
        for (int k = 0; k < d; k++) planes[(size_t)k * n + i] = rows[i * stride + col0 + k];
    SE3_TRY(dst.ensure(planes.size() * sizeof(double)));
    SE3_CUDA(cudaMemcpyAsync(dst.ptr, planes.data(), planes.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    SE3_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int download_planes(se3icp_ctx* c, const DeviceBuf& src, double* rows, size_t n, int d) {
    std::vector<double> planes((size_t)d * n);
    SE3_CUDA(cudaMemcpyAsync(planes.data(), src.ptr, planes.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    SE3_CUDA(cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < n; i++)
        for (int k = 0; k < d; k++) rows[i * d + k] = planes[(size_t)k * n + i];
    return 0;
}

// 3x3 row-major covariances -> 6 symmetric planes
int upload_cov(se3icp_ctx* c, DeviceBuf& dst, const double* cov9, size_t n) {
    static const int map6[6] = {0, 1, 2, 4, 5, 8};
    std::vector<double> planes(6 * n);
    for (size_t i = 0; i < n; i++)
        for (int e = 0; e < 6; e++) planes[(size_t)e * n + i] = cov9[9 * i + map6[e]];
    SE3_TRY(dst.ensure(planes.size() * sizeof(double)));
    SE3_CUDA(cudaMemcpyAsync(dst.ptr, planes.data(), planes.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    SE3_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

int stage_features(se3icp_ctx* c, const double* xyz, size_t n, int k_lrf, int k_nrm, int k_list, bool want_knn) {
    if (!xyz || n == 0 || k_list <= 0) return SE3ICP_ERR_ARG;
    if (k_list > SE3ICP_MAX_KNN) {
        set_last_error("k = %d above SE3ICP_MAX_KNN", k_list);
        return SE3ICP_ERR_UNSUPPORTED;
    }
    SE3_TRY(upload_cloud_and_index(c, 0, xyz, n));
    FeatureArgs fa{};
    fa.k_lrf = k_lrf;
    fa.k_nrm = k_nrm;
    fa.K = k_list;
    fa.q_begin = 0;
    fa.q_end = 0x7fffffff;
    if (k_lrf > 0) {
        SE3_TRY(c->frame[0].ensure(9 * n * sizeof(double)));
        fa.frame = c->frame[0].as<double>();
    }
    if (k_nrm > 0) {
        SE3_TRY(c->nrm[0].ensure(3 * n * sizeof(double)));
        fa.nrm = c->nrm[0].as<double>();
    }
    if (want_knn) {
        SE3_TRY(c->scratch.ensure(n * (size_t)k_list * (sizeof(int) + sizeof(double))));
        fa.knn_d2 = c->scratch.as<double>();
        fa.knn_idx = reinterpret_cast<int*>(c->scratch.as<double>() + n * (size_t)k_list);
    }
    SE3_TRY(launch_knn_features(c->index[0].view, fa, c->stream));
    return 0;
}

void identity_config(RunConfig& cfg, int variant, bool se3) {
    memset(&cfg, 0, sizeof(cfg));
    cfg.entry = se3 ? SE3ICP_RUN_SE3_PURE : SE3ICP_RUN_ICP;
    cfg.variant = variant;
    cfg.max_iter = 1;
    cfg.max_se3_iter = 1;
    cfg.has_se3 = se3;
    cfg.pure = se3;
    cfg.alpha = 1.0;
    cfg.beta = 1.0;
    cfg.mse = 0;
    cfg.mse_switch = 0;
}

int stage_state(se3icp_ctx* c) {
    SE3_TRY(c->state.ensure(sizeof(IterState)));
    SE3_TRY(c->hist.ensure(4 * 256 * sizeof(unsigned int)));
    SE3_TRY(launch_init_state(c->dstate(), c->hist.as<unsigned int>(), c->stream));
    return 0;
}

int stage_corr_alloc(se3icp_ctx* c, size_t n) {
    SE3_TRY(c->corr_idx.ensure(n * sizeof(int)));
    SE3_TRY(c->corr_dist.ensure(n * sizeof(double)));
    SE3_TRY(c->corr_distf.ensure(n * sizeof(float)));
    SE3_TRY(c->keep.ensure(n));
    SE3_TRY(c->repair.ensure(n * sizeof(int)));
    SE3_TRY(c->d2_nd.ensure(n * sizeof(double)));
    SE3_CUDA(cudaMemsetAsync(c->corr_idx.ptr, 0xff, n * sizeof(int), c->stream));
    return 0;
}

}  // namespace

int se3icp_knn(se3icp_ctx* c, const double* xyz, size_t n, int k, int32_t* idx, double* d2) {
    SE3_TRY(check_ctx(c));
    SE3_NOT_PENDING("stage entry point");
    StageScope stage_scope{c};
    if (!idx) return SE3ICP_ERR_ARG;
    SE3_TRY(stage_features(c, xyz, n, 0, 0, k, true));
    const double* dd = c->scratch.as<double>();
    const int* di = reinterpret_cast<const int*>(dd + n * (size_t)k);
    SE3_CUDA(cudaMemcpyAsync(idx, di, n * (size_t)k * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (d2) SE3_CUDA(cudaMemcpyAsync(d2, dd, n * (size_t)k * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    SE3_CUDA(cudaStreamSynchronize(c->stream));
    return SE3ICP_OK;
}

int se3icp_lrf(se3icp_ctx* c, const double* xyz, size_t n, int k, double* frames) {
    SE3_TRY(check_ctx(c));
    SE3_NOT_PENDING("stage entry point");
    StageScope stage_scope{c};
    if (!frames) return SE3ICP_ERR_ARG;
    SE3_TRY(stage_features(c, xyz, n, k, 0, k, false));
    std::vector<double> rows(9 * n);
    SE3_TRY(download_planes(c, c->frame[0], rows.data(), n, 9));
    for (size_t i = 0; i < n; i++) {
        double* F = frames + 16 * i;
        const double* r = &rows[9 * i];  // x-axis, y-axis, z-axis
        for (int col = 0; col < 3; col++)
            for (int row = 0; row < 3; row++) F[4 * row + col] = r[3 * col + row];
        F[3] = xyz[3 * i], F[7] = xyz[3 * i + 1], F[11] = xyz[3 * i + 2];
        F[12] = F[13] = F[14] = 0.0;
        F[15] = 1.0;
    }
    return SE3ICP_OK;
}

int se3icp_shot_lrf(se3icp_ctx* c, const double* xyz, size_t n, double radius, double* frames, int64_t* unresolved_ties) {
    SE3_TRY(check_ctx(c));
    SE3_NOT_PENDING("stage entry point");
    StageScope stage_scope{c};
    if (!xyz || !frames || n == 0 || !(radius > 0.0)) return SE3ICP_ERR_ARG;
    SE3_TRY(upload_cloud_and_index(c, 0, xyz, n));
    SE3_TRY(c->frame[0].ensure(9 * n * sizeof(double)));
    SE3_TRY(c->knn_count.ensure(sizeof(int)));
    SE3_CUDA(cudaMemsetAsync(c->knn_count.ptr, 0, sizeof(int), c->stream));
    SE3_TRY(launch_shot_lrf(c->index[0].view, radius, c->frame[0].as<double>(), c->knn_count.as<int>(), c->stream));
    int unresolved = 0;
    SE3_CUDA(cudaMemcpyAsync(&unresolved, c->knn_count.ptr, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    std::vector<double> rows(9 * n);
    SE3_TRY(download_planes(c, c->frame[0], rows.data(), n, 9));  // synchronises the stream
    if (unresolved_ties) *unresolved_ties = unresolved;
    for (size_t i = 0; i < n; i++) {
        double* F = frames + 16 * i;
        const double* r = &rows[9 * i];  // x-axis, y-axis, z-axis
        for (int col = 0; col < 3; col++)
            for (int row = 0; row < 3; row++) F[4 * row + col] = r[3 * col + row];
        F[3] = xyz[3 * i], F[7] = xyz[3 * i + 1], F[11] = xyz[3 * i + 2];
        F[12] = F[13] = F[14] = 0.0;
        F[15] = 1.0;
    }
    return SE3ICP_OK;
}

int se3icp_normals(se3icp_ctx* c, const double* xyz, size_t n, int k, double* normals) {
    SE3_TRY(check_ctx(c));
    SE3_NOT_PENDING("stage entry point");
    StageScope stage_scope{c};
    if (!normals) return SE3ICP_ERR_ARG;
    SE3_TRY(stage_features(c, xyz, n, 0, k, k, false));
    return download_planes(c, c->nrm[0], normals, n, 3);
}

int se3icp_gicp_cov(se3icp_ctx* c, const double* normals, size_t n, double eps, double* cov) {
    SE3_TRY(check_ctx(c));
    SE3_NOT_PENDING("stage entry point");
    StageScope stage_scope{c};
    if (!normals || !cov || n == 0) return SE3ICP_ERR_ARG;
    SE3_TRY(upload_planes(c, c->nrm[0], normals, n, 3, 3, 0));
    SE3_TRY(c->cov[0].ensure(6 * n * sizeof(double)));
    SE3_TRY(launch_cov_from_normals(c->nrm[0].as<double>(), (int)n, eps, c->cov[0].as<double>(), c->stream));
    std::vector<double> c6(6 * n);
    SE3_TRY(download_planes(c, c->cov[0], c6.data(), n, 6));
    for (size_t i = 0; i < n; i++) {
        const double* s = &c6[6 * i];
        double* o = cov + 9 * i;
        o[0] = s[0], o[1] = s[1], o[2] = s[2];
        o[3] = s[1], o[4] = s[3], o[5] = s[4];
        o[6] = s[2], o[7] = s[4], o[8] = s[5];
    }
    return SE3ICP_OK;
}

int se3icp_nn_se3(se3icp_ctx* c, const double* src_rows, size_t n, const double* tgt_rows, size_t m, int nn_mode,
                  int32_t* idx, double* d2, int64_t* exact_repairs) {
    SE3_TRY(check_ctx(c));
    SE3_NOT_PENDING("stage entry point");
    StageScope stage_scope{c};
    if (!src_rows || !tgt_rows || !idx || n == 0 || m == 0) return SE3ICP_ERR_ARG;
    // target: spatial order from the translation part, rotation planes as given (alpha = beta = 1)
    std::vector<double> xyz(3 * m);
    for (size_t j = 0; j < m; j++)
        for (int k = 0; k < 3; k++) xyz[3 * j + k] = tgt_rows[12 * j + 9 + k];
    SE3_TRY(upload_cloud_and_index(c, 1, xyz.data(), m));
    SE3_TRY(upload_planes(c, c->frame[1], tgt_rows, m, 9, 12, 0));
    SE3_TRY(stage_state(c));
    SE3_TRY(c->se3idx.reserve((int)m, c->index[1].view));
    SE3_TRY(c->se3idx.build(c->index[1].view, c->frame[1].as<double>(), 1.0, 1.0, c->dstate(), c->stream, nullptr));
    // source
    SE3_TRY(upload_planes(c, c->frame[0], src_rows, n, 9, 12, 0));
    SE3_TRY(upload_planes(c, c->scratch, src_rows, n, 3, 12, 9));
    c->n[0] = n;
    SourceView S{};
    S.n = (int)n;
    S.begin = 0;
    S.end = (int)n;
    S.order = nullptr;
    S.x = c->scratch.as<double>();
    S.y = S.x + n;
    S.z = S.x + 2 * n;
    S.frame = c->frame[0].as<double>();
    RunConfig cfg;
    identity_config(cfg, SE3ICP_PT2PT, true);
    c->cfg = cfg;  // target_view() reads beta / with_cf from the context's config
    TargetView T = c->target_view();
    SE3_TRY(stage_corr_alloc(c, n));
    CorrBuffers cb = c->corr_buffers(true);
    switch (nn_mode) {
        case SE3ICP_NN_AUTO:
        case SE3ICP_NN_TREE:
            SE3_TRY(launch_nn_filter(S, T, cfg, c->dstate(), cb, c->stream));  // seeds every query
            SE3_TRY(launch_nn_se3_tree(S, T, cfg, c->dstate(), cb, c->stream));
            break;
        case SE3ICP_NN_BRUTE_F32:
        case SE3ICP_NN_EXACT_F64:
            SE3_TRY(launch_nn_se3_brute(S, T, cfg, c->dstate(), cb, nn_mode == SE3ICP_NN_EXACT_F64, c->stream));
            SE3_TRY(launch_nn_se3_repair(S, T, cfg, c->dstate(), cb, c->stream));
            break;
        default:
            set_last_error("nn_mode %d not available", nn_mode);
            return SE3ICP_ERR_UNSUPPORTED;
    }
    SE3_CUDA(cudaMemcpyAsync(c->h_state, c->dstate(), sizeof(IterState), cudaMemcpyDeviceToHost, c->stream));
    SE3_CUDA(cudaMemcpyAsync(idx, cb.idx, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (d2) SE3_CUDA(cudaMemcpyAsync(d2, cb.d2_nd, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    SE3_CUDA(cudaStreamSynchronize(c->stream));
    if (exact_repairs) *exact_repairs = c->h_state->repair_count;
    return SE3ICP_OK;
}

int se3icp_nn_xyz(se3icp_ctx* c, const double* queries, size_t n, const double* tgt_xyz, size_t m, int32_t* idx,
                  double* d2) {
    SE3_TRY(check_ctx(c));
    SE3_NOT_PENDING("stage entry point");
    StageScope stage_scope{c};
    if (!queries || !tgt_xyz || !idx || n == 0 || m == 0) return SE3ICP_ERR_ARG;
    SE3_TRY(upload_cloud_and_index(c, 1, tgt_xyz, m));
    SE3_TRY(upload_planes(c, c->scratch, queries, n, 3, 3, 0));
    SE3_TRY(stage_state(c));
    c->n[0] = n;
    SourceView S{};
    S.n = (int)n;
    S.begin = 0;
    S.end = (int)n;
    S.order = nullptr;
    S.x = c->scratch.as<double>();
    S.y = S.x + n;
    S.z = S.x + 2 * n;
    TargetView T = c->target_view();
    RunConfig cfg;
    identity_config(cfg, SE3ICP_PT2PT, false);
    SE3_TRY(stage_corr_alloc(c, n));
    CorrBuffers cb = c->corr_buffers(true);
    SE3_TRY(launch_nn_filter(S, T, cfg, c->dstate(), cb, c->stream));  // seeds every query
    SE3_TRY(launch_nn_xyz(S, T, cfg, c->dstate(), cb, c->stream));
    SE3_CUDA(cudaMemcpyAsync(idx, cb.idx, n * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    if (d2) SE3_CUDA(cudaMemcpyAsync(d2, cb.d2_nd, n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    SE3_CUDA(cudaStreamSynchronize(c->stream));
    return SE3ICP_OK;
}

int se3icp_trim(se3icp_ctx* c, const float* dist, size_t n, double overlap, int keep_largest, uint8_t* keep,
                int64_t* n_keep) {
    SE3_TRY(check_ctx(c));
    SE3_NOT_PENDING("stage entry point");
    StageScope stage_scope{c};
    if (!dist || !keep || n == 0) return SE3ICP_ERR_ARG;
    SE3_TRY(stage_state(c));
    SE3_TRY(stage_corr_alloc(c, n));
    SE3_TRY(c->block_eq.ensure((size_t)kReduceBlocks * sizeof(int)));
    SE3_CUDA(cudaMemcpyAsync(c->corr_distf.ptr, dist, n * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    RunConfig cfg;
    identity_config(cfg, SE3ICP_PT2PT, false);
    size_t k = std::min(trimmed_count(n, overlap), n);
    cfg.trim_active = k < n;
    cfg.n_keep_target = (int)k;
    cfg.keep_largest = keep_largest != 0;
    if (n_keep) *n_keep = (int64_t)k;
    if (!cfg.trim_active) {  // pass-through branch of the rejector
        memset(keep, 1, n);
        return SE3ICP_OK;
    }
    if (k == 0) {
        memset(keep, 0, n);
        return SE3ICP_OK;
    }
    // both implementations of the rejection are reachable from here: the single-launch selection the registration loop
    // uses on one GPU (default) and the multi-pass mask kernels of the sharded pair (SE3ICP_TRIM_MULTIPASS=1)
    static const bool multipass = [] {
        const char* e = getenv("SE3ICP_TRIM_MULTIPASS");
        return e && atoi(e) != 0;
    }();
    SE3_TRY(c->thist.ensure((size_t)kTrimHistBins * sizeof(unsigned int)));
    SE3_TRY(c->tcand.ensure(n * sizeof(unsigned long long)));
    SE3_TRY(c->tcount.ensure(kTcountWords * sizeof(unsigned int)));
    CorrBuffers cb = c->corr_buffers(false);
    cb.thist = c->thist.as<unsigned int>();
    cb.tcand = c->tcand.as<unsigned long long>();
    cb.tcount = c->tcount.as<unsigned int>();
    if (multipass)
        SE3_TRY(launch_trim(cfg, c->dstate(), cb, (int)n, c->hist.as<unsigned int>(), c->block_eq.as<int>(), c->stream));
    else
        SE3_TRY(launch_trim_stage(cfg, c->dstate(), cb, (int)n, c->stream));
    SE3_CUDA(cudaMemcpyAsync(keep, c->keep.ptr, n, cudaMemcpyDeviceToHost, c->stream));
    SE3_CUDA(cudaStreamSynchronize(c->stream));
    return SE3ICP_OK;
}

namespace {

// shared body of the three reduce entry points: uploads, runs reduce (+ optionally solve), returns sums
int stage_reduce(se3icp_ctx* c, int variant, const double* src, const double* src_cov, size_t n, const double* tgt,
                 const double* tgt_nrm, const double* tgt_cov, size_t m, const int32_t* corr_tgt, const double* conf_src,
                 const double* conf_tgt, double* out27, double* T_out) {
    if (!src || !tgt || !corr_tgt || n == 0 || m == 0) return SE3ICP_ERR_ARG;
    SE3_TRY(upload_cloud_and_index(c, 1, tgt, m));
    SE3_TRY(upload_planes(c, c->scratch, src, n, 3, 3, 0));
    if (tgt_nrm) SE3_TRY(upload_planes(c, c->nrm[1], tgt_nrm, m, 3, 3, 0));
    if (tgt_cov) SE3_TRY(upload_cov(c, c->cov[1], tgt_cov, m));
    if (src_cov) SE3_TRY(upload_cov(c, c->cov[0], src_cov, n));
    if (conf_src) SE3_TRY(upload_planes(c, c->conf[0], conf_src, n, 1, 1, 0));
    if (conf_tgt) SE3_TRY(upload_planes(c, c->conf[1], conf_tgt, m, 1, 1, 0));
    SE3_TRY(stage_state(c));
    SE3_TRY(stage_corr_alloc(c, n));
    SE3_CUDA(cudaMemcpyAsync(c->corr_idx.ptr, corr_tgt, n * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    SE3_CUDA(cudaMemsetAsync(c->corr_distf.ptr, 0, n * sizeof(float), c->stream));
    SE3_TRY(c->partials.ensure((size_t)kReduceBlocks * kReducePartials * sizeof(double)));
    c->n[0] = n;
    SourceView S{};
    S.n = (int)n;
    S.begin = 0;
    S.end = (int)n;
    S.order = nullptr;
    S.x = c->scratch.as<double>();
    S.y = S.x + n;
    S.z = S.x + 2 * n;
    S.cov = c->cov[0].as<double>();
    S.conf = c->conf[0].as<double>();
    TargetView T = c->target_view();
    RunConfig cfg;
    identity_config(cfg, variant, false);
    cfg.with_cf = (conf_src && conf_tgt) ? 1 : 0;
    SE3_TRY(launch_reduce(S, T, cfg, c->dstate(), c->corr_buffers(false), c->partials.as<double>(), SolveFusion{}, c->stream));
    if (out27) {
        std::vector<double> part((size_t)kReduceBlocks * kReducePartials);
        SE3_CUDA(cudaMemcpyAsync(part.data(), c->partials.ptr, part.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
        SE3_CUDA(cudaStreamSynchronize(c->stream));
        for (int k = 0; k < 27; k++) {
            double s = 0.0;
            for (int b = 0; b < kReduceBlocks; b++) s += part[(size_t)b * kReducePartials + k];
            out27[k] = s;
        }
    }
    if (T_out) {
        SE3_TRY(launch_solve_update(cfg, c->dstate(), c->partials.as<double>(), kReduceBlocks, nullptr, c->hist.as<unsigned int>(), 0, c->stream));
        SE3_CUDA(cudaMemcpyAsync(c->h_state, c->dstate(), sizeof(IterState), cudaMemcpyDeviceToHost, c->stream));
        SE3_CUDA(cudaStreamSynchronize(c->stream));
        memcpy(T_out, c->h_state->T_i, 16 * sizeof(double));
    }
    return SE3ICP_OK;
}

}  // namespace

int se3icp_reduce_pt2pt(se3icp_ctx* c, const double* src, size_t n, const double* tgt, size_t m, const int32_t* corr_tgt,
                        double* T_out) {
    SE3_TRY(check_ctx(c));
    SE3_NOT_PENDING("stage entry point");
    StageScope stage_scope{c};
    if (!T_out) return SE3ICP_ERR_ARG;
    return stage_reduce(c, SE3ICP_PT2PT, src, nullptr, n, tgt, nullptr, nullptr, m, corr_tgt, nullptr, nullptr, nullptr, T_out);
}

int se3icp_reduce_pt2pl(se3icp_ctx* c, const double* src, size_t n, const double* tgt, const double* tgt_normals, size_t m,
                        const int32_t* corr_tgt, double* out27) {
    SE3_TRY(check_ctx(c));
    SE3_NOT_PENDING("stage entry point");
    StageScope stage_scope{c};
    if (!tgt_normals || !out27) return SE3ICP_ERR_ARG;
    return stage_reduce(c, SE3ICP_PT2PL, src, nullptr, n, tgt, tgt_normals, nullptr, m, corr_tgt, nullptr, nullptr, out27,
                        nullptr);
}

int se3icp_reduce_gicp(se3icp_ctx* c, const double* src, const double* src_cov, size_t n, const double* tgt,
                       const double* tgt_cov, size_t m, const int32_t* corr_tgt, const double* conf_src,
                       const double* conf_tgt, double* out27) {
    SE3_TRY(check_ctx(c));
    SE3_NOT_PENDING("stage entry point");
    StageScope stage_scope{c};
    if (!src_cov || !tgt_cov || !out27) return SE3ICP_ERR_ARG;
    return stage_reduce(c, SE3ICP_GICP, src, src_cov, n, tgt, nullptr, tgt_cov, m, corr_tgt, conf_src, conf_tgt, out27,
                        nullptr);
}

int se3icp_solve(se3icp_ctx* c, const double* in27, double* T_out) {
    SE3_TRY(check_ctx(c));
    SE3_NOT_PENDING("stage entry point");
    StageScope stage_scope{c};
    if (!in27 || !T_out) return SE3ICP_ERR_ARG;
    SE3_TRY(stage_state(c));
    SE3_TRY(c->partials.ensure((size_t)kReduceBlocks * kReducePartials * sizeof(double)));
    std::vector<double> part((size_t)kReduceBlocks * kReducePartials, 0.0);
    for (int k = 0; k < 27; k++) part[k] = in27[k];
    part[28] = 1.0;  // one correspondence, so the solve is not skipped
    SE3_CUDA(cudaMemcpyAsync(c->partials.ptr, part.data(), part.size() * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    RunConfig cfg;
    identity_config(cfg, SE3ICP_PT2PL, false);
    SE3_TRY(launch_solve_update(cfg, c->dstate(), c->partials.as<double>(), kReduceBlocks, nullptr, c->hist.as<unsigned int>(), 0, c->stream));
    SE3_CUDA(cudaMemcpyAsync(c->h_state, c->dstate(), sizeof(IterState), cudaMemcpyDeviceToHost, c->stream));
    SE3_CUDA(cudaStreamSynchronize(c->stream));
    memcpy(T_out, c->h_state->T_i, 16 * sizeof(double));
    return SE3ICP_OK;
}

}  // extern "C"
