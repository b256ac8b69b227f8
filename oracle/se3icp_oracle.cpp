/*
 * se3icp_oracle.cpp — CPU oracle for the SE(3)-ICP registration path.
 *
 * TEST INFRASTRUCTURE ONLY (see se3icp_oracle.h).  Pinned against the
 * reference's own source through oracle/_ref (tests/test_reference_build.py);
 * PARITY UNPINNED at the Open3D / PCL / Eigen boundaries: those libraries are
 * not vendored in the reference and are absent from this image, so their
 * behaviour is restated here from their published algorithms:
 *   - Open3D 0.19.0 @1868f4332: KDTreeFlann (nanoflann, exact L2 kNN, results
 *     ascending), PointCloud::{GetCenter,Translate,Scale,Transform,
 *     EstimateNormals}, TransformationEstimation{PointToPoint (Eigen::umeyama,
 *     no scaling), PointToPlane, ForGeneralizedICP}, ComputeJTJandJTr,
 *     SolveJacobianSystemAndObtainExtrinsicMatrix (LDLT; x -> Rz*Ry*Rx, t).
 *   - PCL 1.14: registration::CorrespondenceRejectorTrimmed.
 *   - Eigen >= 3.3: SelfAdjointEigenSolver<Matrix3d>, inverse().sqrt().
 * Every function cites the reference lines (relative to /root/reference) that
 * it follows.  All arithmetic is FP64 except the stored correspondence distance
 * (float), exactly as in the reference.  Built with -ffp-contract=off so that
 * squared distances are plain mul/add chains the CUDA path can reproduce bit
 * for bit.
 *
 * Deliberate, documented deviations (none changes a result beyond a tie):
 *   - equal-distance ties in every nearest-neighbour query resolve to the
 *     smallest point index (nanoflann: first found in traversal order);
 *   - normals/eigenvectors use a cyclic Jacobi solver (Eigen: tridiagonal QL,
 *     Open3D: analytic FastEigen3x3) — same eigenvectors up to sign.
 */
#include "se3icp_oracle.h"

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <numeric>
#include <vector>

#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

using Clock = std::chrono::high_resolution_clock;
inline double ms_since(Clock::time_point t0) {
    return std::chrono::duration_cast<std::chrono::nanoseconds>(Clock::now() - t0).count() / 1e6;
}

// ----------------------------------------------------------------------------------------------
// small fixed-size linear algebra (row-major)
// ----------------------------------------------------------------------------------------------
struct V3 {
    double x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double norm(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 load3(const double* p) { return {p[0], p[1], p[2]}; }

struct M3 {
    double a[9];
    double& operator()(int r, int c) { return a[3 * r + c]; }
    double operator()(int r, int c) const { return a[3 * r + c]; }
};
inline M3 m3_zero() {
    M3 m;
    for (double& v : m.a) v = 0.0;
    return m;
}
inline M3 m3_identity() {
    M3 m = m3_zero();
    m(0, 0) = m(1, 1) = m(2, 2) = 1.0;
    return m;
}
inline M3 m3_mul(const M3& A, const M3& B) {
    M3 C = m3_zero();
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            for (int k = 0; k < 3; k++) C(i, j) += A(i, k) * B(k, j);
    return C;
}
inline M3 m3_transpose(const M3& A) {
    M3 T;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) T(i, j) = A(j, i);
    return T;
}
inline V3 m3_mulv(const M3& A, V3 v) {
    return {A(0, 0) * v.x + A(0, 1) * v.y + A(0, 2) * v.z, A(1, 0) * v.x + A(1, 1) * v.y + A(1, 2) * v.z,
            A(2, 0) * v.x + A(2, 1) * v.y + A(2, 2) * v.z};
}
inline double m3_det(const M3& A) {
    return A(0, 0) * (A(1, 1) * A(2, 2) - A(1, 2) * A(2, 1)) - A(0, 1) * (A(1, 0) * A(2, 2) - A(1, 2) * A(2, 0)) +
           A(0, 2) * (A(1, 0) * A(2, 1) - A(1, 1) * A(2, 0));
}
inline M3 m3_skew(V3 v) {  // Open3D utility::SkewMatrix
    M3 S = m3_zero();
    S(0, 1) = -v.z;
    S(0, 2) = v.y;
    S(1, 0) = v.z;
    S(1, 2) = -v.x;
    S(2, 0) = -v.y;
    S(2, 1) = v.x;
    return S;
}

struct M4 {
    double a[16];
    double& operator()(int r, int c) { return a[4 * r + c]; }
    double operator()(int r, int c) const { return a[4 * r + c]; }
};
inline M4 m4_identity() {
    M4 m;
    for (int i = 0; i < 16; i++) m.a[i] = (i % 5 == 0) ? 1.0 : 0.0;
    return m;
}
inline M4 m4_mul(const M4& A, const M4& B) {
    M4 C;
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            double s = 0.0;
            for (int k = 0; k < 4; k++) s += A(i, k) * B(k, j);
            C(i, j) = s;
        }
    return C;
}
inline double m4_diff_fro(const M4& A, const M4& B) {
    double s = 0.0;
    for (int i = 0; i < 16; i++) s += (A.a[i] - B.a[i]) * (A.a[i] - B.a[i]);
    return std::sqrt(s);
}

// Symmetric 3x3 eigen-decomposition by cyclic Jacobi rotations; eigenvalues ascending, unit
// eigenvectors in the columns of V.  Stands in for Eigen::SelfAdjointEigenSolver<Matrix3d>
// (reference .cpp:275-281) and Open3D's FastEigen3x3 (EstimateNormals).
void eig3_sym(const M3& Ain, double evals[3], M3& V) {
    M3 A = Ain;
    V = m3_identity();
    for (int sweep = 0; sweep < 64; sweep++) {
        double off = A(0, 1) * A(0, 1) + A(0, 2) * A(0, 2) + A(1, 2) * A(1, 2);
        double diag = A(0, 0) * A(0, 0) + A(1, 1) * A(1, 1) + A(2, 2) * A(2, 2);
        if (off == 0.0 || off <= 1e-32 * diag) break;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                double apq = A(p, q);
                if (apq == 0.0) continue;
                double theta = (A(q, q) - A(p, p)) / (2.0 * apq);
                double t = (theta >= 0.0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                // A <- J^T A J with J the rotation in the (p,q) plane
                for (int k = 0; k < 3; k++) {
                    double akp = A(k, p), akq = A(k, q);
                    A(k, p) = c * akp - s * akq;
                    A(k, q) = s * akp + c * akq;
                }
                for (int k = 0; k < 3; k++) {
                    double apk = A(p, k), aqk = A(q, k);
                    A(p, k) = c * apk - s * aqk;
                    A(q, k) = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; k++) {
                    double vkp = V(k, p), vkq = V(k, q);
                    V(k, p) = c * vkp - s * vkq;
                    V(k, q) = s * vkp + c * vkq;
                }
            }
    }
    int order[3] = {0, 1, 2};
    double d[3] = {A(0, 0), A(1, 1), A(2, 2)};
    std::sort(order, order + 3, [&](int i, int j) { return d[i] < d[j]; });
    M3 Vs;
    for (int c = 0; c < 3; c++) {
        evals[c] = d[order[c]];
        for (int r = 0; r < 3; r++) Vs(r, c) = V(r, order[c]);
    }
    V = Vs;
}

// ----------------------------------------------------------------------------------------------
// exact kd-tree (any dimension), stands in for open3d::geometry::KDTreeFlann / nanoflann
// (reference call sites .cpp:253,407,458,482,586-587,626).  Leaf size 15 as nanoflann's
// KDTreeEigenMatrixAdaptor is configured by Open3D.
// ----------------------------------------------------------------------------------------------
struct Neighbor {
    double d2;
    int idx;
};
inline bool nb_less(const Neighbor& a, const Neighbor& b) { return a.d2 < b.d2 || (a.d2 == b.d2 && a.idx < b.idx); }

class KDTree {
public:
    void build(const double* data, size_t n, int dim) {
        dim_ = dim;
        n_ = n;
        pts_.assign(data, data + n * (size_t)dim);
        perm_.resize(n);
        std::iota(perm_.begin(), perm_.end(), 0);
        nodes_.clear();
        nodes_.reserve(2 * (n / kLeaf + 1));
        lo_.assign(dim, std::numeric_limits<double>::infinity());
        hi_.assign(dim, -std::numeric_limits<double>::infinity());
        for (size_t i = 0; i < n; i++)
            for (int d = 0; d < dim; d++) {
                lo_[d] = std::min(lo_[d], pts_[i * dim + d]);
                hi_[d] = std::max(hi_[d], pts_[i * dim + d]);
            }
        if (n > 0) build_rec(0, (int)n);
    }
    size_t size() const { return n_; }
    int dim() const { return dim_; }
    const double* point(int i) const { return &pts_[(size_t)i * dim_]; }

    // k nearest, ascending by (d2, idx); returns the number found (min(k, n)).
    int knn(const double* q, int k, Neighbor* out) const {
        if (n_ == 0 || k <= 0) return 0;
        Search s;
        s.q = q;
        s.k = std::min<size_t>(k, n_);
        s.heap = out;
        s.count = 0;
        std::vector<double> side(dim_, 0.0);
        double mind = 0.0;
        for (int d = 0; d < dim_; d++) {
            double v = q[d];
            if (v < lo_[d]) side[d] = (v - lo_[d]) * (v - lo_[d]);
            if (v > hi_[d]) side[d] = (v - hi_[d]) * (v - hi_[d]);
            mind += side[d];
        }
        search_rec(0, mind, side.data(), s);
        std::sort_heap(out, out + s.count, nb_less);
        return (int)s.count;
    }

    // every point with d2 < r2 (nanoflann's RadiusResultSet::addPoint: strict), ascending by (d2, idx) — Open3D
    // KDTreeFlann::SearchRadius (sorted = true); call site reference .cpp:136
    void radius(const double* q, double r2, std::vector<Neighbor>& out) const {
        out.clear();
        if (n_ == 0) return;
        std::vector<double> side(dim_, 0.0);
        double mind = 0.0;
        for (int d = 0; d < dim_; d++) {
            double v = q[d];
            if (v < lo_[d]) side[d] = (v - lo_[d]) * (v - lo_[d]);
            if (v > hi_[d]) side[d] = (v - hi_[d]) * (v - hi_[d]);
            mind += side[d];
        }
        radius_rec(0, mind, side.data(), q, r2, out);
        std::sort(out.begin(), out.end(), nb_less);
    }

private:
    static constexpr int kLeaf = 15;
    struct Node {
        int left = -1, right = -1;  // children, or [begin,end) into perm_ for leaves
        int split_dim = -1;         // -1 marks a leaf
        double split_lo = 0, split_hi = 0;
    };
    struct Search {
        const double* q;
        size_t k;
        Neighbor* heap;  // max-heap under nb_less while searching
        size_t count;
    };
    int dim_ = 0;
    size_t n_ = 0;
    std::vector<double> pts_, lo_, hi_;
    std::vector<int> perm_;
    std::vector<Node> nodes_;

    int build_rec(int begin, int end) {
        int id = (int)nodes_.size();
        nodes_.emplace_back();
        if (end - begin <= kLeaf) {
            nodes_[id].left = begin;
            nodes_[id].right = end;
            return id;
        }
        // split on the dimension with the largest spread among the points of this cell, at the median
        int best_dim = 0;
        double best_span = -1.0;
        for (int d = 0; d < dim_; d++) {
            double lo = std::numeric_limits<double>::infinity(), hi = -lo;
            for (int i = begin; i < end; i++) {
                double v = pts_[(size_t)perm_[i] * dim_ + d];
                lo = std::min(lo, v);
                hi = std::max(hi, v);
            }
            if (hi - lo > best_span) {
                best_span = hi - lo;
                best_dim = d;
            }
        }
        int mid = begin + (end - begin) / 2;
        std::nth_element(perm_.begin() + begin, perm_.begin() + mid, perm_.begin() + end, [&](int a, int b) {
            return pts_[(size_t)a * dim_ + best_dim] < pts_[(size_t)b * dim_ + best_dim];
        });
        double left_hi = -std::numeric_limits<double>::infinity(), right_lo = std::numeric_limits<double>::infinity();
        for (int i = begin; i < mid; i++) left_hi = std::max(left_hi, pts_[(size_t)perm_[i] * dim_ + best_dim]);
        for (int i = mid; i < end; i++) right_lo = std::min(right_lo, pts_[(size_t)perm_[i] * dim_ + best_dim]);
        nodes_[id].split_dim = best_dim;
        nodes_[id].split_lo = left_hi;
        nodes_[id].split_hi = right_lo;
        int l = build_rec(begin, mid);
        int r = build_rec(mid, end);
        nodes_[id].left = l;
        nodes_[id].right = r;
        return id;
    }

    void radius_rec(int id, double mind, double* side, const double* q, double r2, std::vector<Neighbor>& out) const {
        const Node& nd = nodes_[id];
        if (nd.split_dim < 0) {
            for (int i = nd.left; i < nd.right; i++) {
                int pi = perm_[i];
                const double* p = &pts_[(size_t)pi * dim_];
                double d2 = 0.0;
                for (int d = 0; d < dim_; d++) {
                    double df = q[d] - p[d];
                    d2 += df * df;
                }
                if (d2 < r2) out.push_back(Neighbor{d2, pi});
            }
            return;
        }
        int d = nd.split_dim;
        double dl = q[d] - nd.split_lo, dh = q[d] - nd.split_hi;
        int near = dl + dh < 0.0 ? nd.left : nd.right, far = dl + dh < 0.0 ? nd.right : nd.left;
        double cut = dl + dh < 0.0 ? dh * dh : dl * dl;
        radius_rec(near, mind, side, q, r2, out);
        double saved = side[d];
        double far_mind = mind + cut - saved;
        if (far_mind * (1.0 - 1e-12) <= r2) {
            side[d] = cut;
            radius_rec(far, far_mind, side, q, r2, out);
            side[d] = saved;
        }
    }

    inline double worst(const Search& s) const {
        return s.count < s.k ? std::numeric_limits<double>::infinity() : s.heap[0].d2;
    }

    void search_rec(int id, double mind, double* side, Search& s) const {
        const Node& nd = nodes_[id];
        if (nd.split_dim < 0) {
            for (int i = nd.left; i < nd.right; i++) {
                int pi = perm_[i];
                const double* p = &pts_[(size_t)pi * dim_];
                double d2 = 0.0;
                for (int d = 0; d < dim_; d++) {
                    double df = s.q[d] - p[d];
                    d2 += df * df;
                }
                Neighbor cand{d2, pi};
                if (s.count < s.k) {
                    s.heap[s.count++] = cand;
                    std::push_heap(s.heap, s.heap + s.count, nb_less);
                } else if (nb_less(cand, s.heap[0])) {
                    std::pop_heap(s.heap, s.heap + s.count, nb_less);
                    s.heap[s.count - 1] = cand;
                    std::push_heap(s.heap, s.heap + s.count, nb_less);
                }
            }
            return;
        }
        int d = nd.split_dim;
        double v = s.q[d];
        double dl = v - nd.split_lo, dh = v - nd.split_hi;
        int near, far;
        double cut;
        if (dl + dh < 0.0) {
            near = nd.left;
            far = nd.right;
            cut = dh * dh;
        } else {
            near = nd.right;
            far = nd.left;
            cut = dl * dl;
        }
        search_rec(near, mind, side, s);
        double saved = side[d];
        double far_mind = mind + cut - saved;
        // keep a tiny safety margin so rounding in the incremental bound can never hide an exact tie
        if (far_mind * (1.0 - 1e-12) <= worst(s)) {
            side[d] = cut;
            search_rec(far, far_mind, side, s);
            side[d] = saved;
        }
    }
};

// ----------------------------------------------------------------------------------------------
// point cloud container mirroring the three attributes the reference path touches
// ----------------------------------------------------------------------------------------------
struct Cloud {
    std::vector<double> pts;      // n*3
    std::vector<double> normals;  // n*3 or empty
    std::vector<double> covs;     // n*9 or empty
    size_t size() const { return pts.size() / 3; }
    V3 p(size_t i) const { return load3(&pts[3 * i]); }
};

// Open3D PointCloud::GetCenter — arithmetic mean (reference .cpp:568-569).
V3 cloud_center(const Cloud& c) {
    V3 s{0, 0, 0};
    size_t n = c.size();
    for (size_t i = 0; i < n; i++) s = s + c.p(i);
    if (n == 0) return s;
    return (1.0 / (double)n) * s;  // Open3D divides the accumulated sum by the count
}

// reference .cpp:112-119
double largest_distance_from(V3 ref, const Cloud& c) {
    double best = -1.0;
    for (size_t i = 0; i < c.size(); i++) best = std::max(best, norm(c.p(i) - ref));
    return best;
}

void cloud_translate(Cloud& c, V3 t) {
    for (size_t i = 0; i < c.size(); i++) {
        c.pts[3 * i] += t.x;
        c.pts[3 * i + 1] += t.y;
        c.pts[3 * i + 2] += t.z;
    }
}
// Open3D Scale(s, center = 0): p = (p - 0) * s + 0
void cloud_scale(Cloud& c, double s) {
    for (double& v : c.pts) v *= s;
}

// Open3D PointCloud::Transform: points (homogeneous divide), normals (R n), covariances (R C R^T).
void cloud_transform(Cloud& c, const M4& T) {
    size_t n = c.size();
    M3 R;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) R(i, j) = T(i, j);
    M3 Rt = m3_transpose(R);
    bool hn = !c.normals.empty(), hc = !c.covs.empty();
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; i++) {
        double* p = &c.pts[3 * i];
        double x = T(0, 0) * p[0] + T(0, 1) * p[1] + T(0, 2) * p[2] + T(0, 3);
        double y = T(1, 0) * p[0] + T(1, 1) * p[1] + T(1, 2) * p[2] + T(1, 3);
        double z = T(2, 0) * p[0] + T(2, 1) * p[1] + T(2, 2) * p[2] + T(2, 3);
        double w = T(3, 0) * p[0] + T(3, 1) * p[1] + T(3, 2) * p[2] + T(3, 3);
        p[0] = x / w;
        p[1] = y / w;
        p[2] = z / w;
        if (hn) {
            V3 nn = m3_mulv(R, load3(&c.normals[3 * i]));
            c.normals[3 * i] = nn.x;
            c.normals[3 * i + 1] = nn.y;
            c.normals[3 * i + 2] = nn.z;
        }
        if (hc) {
            M3 C;
            std::memcpy(C.a, &c.covs[9 * i], sizeof(C.a));
            M3 RC = m3_mul(m3_mul(R, C), Rt);
            std::memcpy(&c.covs[9 * i], RC.a, sizeof(C.a));
        }
    }
}

// ----------------------------------------------------------------------------------------------
// TOLDI local reference frame, kNN variant — reference .cpp:241-316 (single), :318-331 (loop)
// ----------------------------------------------------------------------------------------------
M4 toldi_frame(const Cloud& cloud, const KDTree& tree, V3 center, int knn_pts, std::vector<Neighbor>& nb) {
    nb.resize(knn_pts);
    double q[3] = {center.x, center.y, center.z};
    int cnt = tree.knn(q, knn_pts, nb.data());  // .cpp:253, ascending
    // .cpp:256 distance to the farthest returned neighbour
    double radius = norm(center - cloud.p(nb[cnt - 1].idx));
    // .cpp:259-265 centroid of neighbours 1 .. cnt/3-1, divided by cnt/3
    V3 centroid{0, 0, 0};
    int rz = cnt / 3;
    for (int i = 1; i < rz; i++) centroid = centroid + cloud.p(nb[i].idx);
    centroid = (1.0 / (double)rz) * centroid;
    // .cpp:268-272 un-normalised scatter of neighbours 1 .. rz about that centroid
    M3 cov = m3_zero();
    for (int i = 1; i < rz + 1; i++) {
        V3 d = cloud.p(nb[i].idx) - centroid;
        double dv[3] = {d.x, d.y, d.z};
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) cov(r, c) += dv[r] * dv[c];
    }
    // .cpp:275-281 normal = eigenvector of the smallest eigenvalue
    double ev[3];
    M3 V;
    eig3_sym(cov, ev, V);
    V3 nrm{V(0, 0), V(1, 0), V(2, 0)};
    // .cpp:286-297 one pass over all neighbours except index 0
    V3 acc{0, 0, 0}, accw{0, 0, 0};
    for (int i = 1; i < cnt; i++) {
        V3 a = cloud.p(nb[i].idx) - center;
        acc = acc + a;
        double nd = dot(nrm, a);
        double an = norm(a);
        double w1 = (radius - an) * (radius - an);
        double w2 = nd * nd;
        accw = accw + (w1 * w2) * a;
    }
    if (dot(nrm, acc) < 0.0) nrm = -1.0 * nrm;  // .cpp:298
    V3 z = nrm;
    V3 x = accw - dot(accw, z) * z;  // .cpp:302 Gram-Schmidt
    x = (1.0 / norm(x)) * x;         // .cpp:303
    V3 y = cross(z, x);              // .cpp:306
    M4 F = m4_identity();
    F(0, 0) = x.x, F(1, 0) = x.y, F(2, 0) = x.z;
    F(0, 1) = y.x, F(1, 1) = y.y, F(2, 1) = y.z;
    F(0, 2) = z.x, F(1, 2) = z.y, F(2, 2) = z.z;
    F(0, 3) = center.x, F(1, 3) = center.y, F(2, 3) = center.z;
    return F;
}

void toldi_all(const Cloud& cloud, const KDTree& tree, int knn_pts, std::vector<M4>& frames) {
    size_t n = cloud.size();
    frames.resize(n);
#pragma omp parallel
    {
        std::vector<Neighbor> nb;
#pragma omp for schedule(dynamic, 64)
        for (long i = 0; i < (long)n; i++) frames[i] = toldi_frame(cloud, tree, cloud.p(i), knn_pts, nb);
    }
}

// ----------------------------------------------------------------------------------------------
// SHOT local reference frame, radius variant — reference .cpp:121-224 (single), :226-239 (loop).  The reference keeps it
// as an alternative to TOLDI (its calls are commented out at .cpp:593-594,812-813; `lrf_radius_` .cpp:340).
//   support   = points with d2 < radius^2, ascending; entry 0 (the centre itself) is skipped              .cpp:136,151
//   M         = sum (radius - d_i) a_i a_i^T / sum (radius - d_i),  a_i = p_i - centre                     .cpp:151-157
//   x+, z+    = eigenvectors of the largest / smallest eigenvalue                                            .cpp:169-170
//   sign      = majority of a_i . v >= 0; on an exact tie the 5 neighbours around the median distance vote   .cpp:172-214
//   y         = z x x                                                                                        .cpp:216
// Fewer than 5 support points: the reference only prints a warning and then divides 0 / 0 or reads diff_vectors out of
// range on a tie; that case is not defined there, and this restatement (like the CUDA path) returns the identity rotation.
// ----------------------------------------------------------------------------------------------
M4 shot_frame(const Cloud& cloud, const KDTree& tree, int index_center, double radius, std::vector<Neighbor>& nb,
              std::vector<V3>& diff) {
    V3 center = cloud.p(index_center);
    double q[3] = {center.x, center.y, center.z};
    tree.radius(q, radius * radius, nb);
    M4 F = m4_identity();
    F(0, 3) = center.x, F(1, 3) = center.y, F(2, 3) = center.z;
    int n_considered = (int)nb.size() - 1;
    if (n_considered < 5) return F;
    M3 cov = m3_zero();
    double total = 0.0;
    diff.clear();
    for (size_t i = 1; i < nb.size(); i++) {
        double w = radius - std::sqrt(nb[i].d2);
        V3 a = cloud.p(nb[i].idx) - center;
        diff.push_back(a);
        double av[3] = {a.x, a.y, a.z};
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) cov(r, c) += w * av[r] * av[c];
        total += w;
    }
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) cov(r, c) /= total;
    double ev[3];
    M3 V;
    eig3_sym(cov, ev, V);
    V3 axis[2] = {{V(0, 2), V(1, 2), V(2, 2)}, {V(0, 0), V(1, 0), V(2, 0)}};  // x+ (largest), z+ (smallest)
    for (int k = 0; k < 2; k++) {
        int plus = 0;
        for (const V3& a : diff)
            if (dot(a, axis[k]) >= 0.0) plus++;
        int s = 2 * plus - n_considered;
        if (s == 0) {
            const int points = 5, median = n_considered / 2;
            for (int i = -points / 2; i <= points / 2; i++)
                if (dot(diff[median - i], axis[k]) >= 0.0) s++;
            if (s < points / 2 + 1) axis[k] = -1.0 * axis[k];
        } else if (s < 0) {
            axis[k] = -1.0 * axis[k];
        }
    }
    V3 x = axis[0], z = axis[1], y = cross(z, x);
    F(0, 0) = x.x, F(1, 0) = x.y, F(2, 0) = x.z;
    F(0, 1) = y.x, F(1, 1) = y.y, F(2, 1) = y.z;
    F(0, 2) = z.x, F(1, 2) = z.y, F(2, 2) = z.z;
    return F;
}

void shot_all(const Cloud& cloud, const KDTree& tree, double radius, std::vector<M4>& frames) {
    size_t n = cloud.size();
    frames.resize(n);
#pragma omp parallel
    {
        std::vector<Neighbor> nb;
        std::vector<V3> diff;
#pragma omp for schedule(dynamic, 64)
        for (long i = 0; i < (long)n; i++) frames[i] = shot_frame(cloud, tree, (int)i, radius, nb, diff);
    }
}

// ----------------------------------------------------------------------------------------------
// Open3D PointCloud::EstimateNormals(KDTreeSearchParamKNN(k)) — reference .cpp:43,494,643.
// kNN including the query point itself; covariance from cumulants E[xx^T] - E[x]E[x]^T; normal =
// eigenvector of the smallest eigenvalue; no orientation; zero-norm fallback (0,0,1).
// ----------------------------------------------------------------------------------------------
void estimate_normals(Cloud& cloud, const KDTree& tree, int k) {
    size_t n = cloud.size();
    cloud.normals.assign(3 * n, 0.0);
#pragma omp parallel
    {
        std::vector<Neighbor> nb(k);
#pragma omp for schedule(dynamic, 64)
        for (long i = 0; i < (long)n; i++) {
            int cnt = tree.knn(&cloud.pts[3 * i], k, nb.data());
            V3 nrm{0, 0, 1};
            if (cnt >= 3) {
                double cum[9] = {0};
                for (int j = 0; j < cnt; j++) {
                    V3 p = cloud.p(nb[j].idx);
                    cum[0] += p.x, cum[1] += p.y, cum[2] += p.z;
                    cum[3] += p.x * p.x, cum[4] += p.x * p.y, cum[5] += p.x * p.z;
                    cum[6] += p.y * p.y, cum[7] += p.y * p.z, cum[8] += p.z * p.z;
                }
                for (double& c : cum) c /= (double)cnt;
                M3 C;
                C(0, 0) = cum[3] - cum[0] * cum[0];
                C(1, 1) = cum[6] - cum[1] * cum[1];
                C(2, 2) = cum[8] - cum[2] * cum[2];
                C(0, 1) = C(1, 0) = cum[4] - cum[0] * cum[1];
                C(0, 2) = C(2, 0) = cum[5] - cum[0] * cum[2];
                C(1, 2) = C(2, 1) = cum[7] - cum[1] * cum[2];
                double ev[3];
                M3 V;
                eig3_sym(C, ev, V);
                nrm = {V(0, 0), V(1, 0), V(2, 0)};
                if (norm(nrm) == 0.0) nrm = {0, 0, 1};
            }
            cloud.normals[3 * i] = nrm.x;
            cloud.normals[3 * i + 1] = nrm.y;
            cloud.normals[3 * i + 2] = nrm.z;
        }
    }
}

// reference .cpp:4-14 GetRotationFromE1ToX, including the c < -0.99 -> Identity quirk
M3 rotation_e1_to(V3 x) {
    V3 e1{1, 0, 0};
    V3 v = cross(e1, x);
    double c = dot(e1, x);
    if (c < -0.99) return m3_identity();
    M3 sv = m3_skew(v);
    double factor = 1.0 / (1.0 + c);
    M3 sv2 = m3_mul(sv, sv);
    M3 R = m3_identity();
    for (int i = 0; i < 9; i++) R.a[i] += sv.a[i] + sv2.a[i] * factor;
    return R;
}

// reference .cpp:45-51 covariance = Rx diag(eps,1,1) Rx^T
void gicp_covariances_from_normals(const std::vector<double>& normals, double eps, std::vector<double>& covs) {
    size_t n = normals.size() / 3;
    covs.resize(9 * n);
    M3 C = m3_zero();
    C(0, 0) = eps, C(1, 1) = 1.0, C(2, 2) = 1.0;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; i++) {
        M3 Rx = rotation_e1_to(load3(&normals[3 * i]));
        M3 out = m3_mul(m3_mul(Rx, C), m3_transpose(Rx));
        std::memcpy(&covs[9 * i], out.a, sizeof(out.a));
    }
}

// reference .cpp:33-52 InitializePointCloudForGeneralizedICP_modified
void init_gicp(Cloud& cloud, const KDTree& tree, int k, double eps) {
    if (!cloud.covs.empty()) return;
    if (cloud.normals.empty()) estimate_normals(cloud, tree, k);
    gicp_covariances_from_normals(cloud.normals, eps, cloud.covs);
}

// reference .cpp:16-30 (note p1*min_depth is not squared — kept as written)
double lounge_point_confidence(V3 v) {
    double depth = v.z;
    double p1 = 0.002203, p2 = -0.001028, p3 = 0.0005351, min_depth = 0.4;
    double error = p1 * depth * depth + p2 * depth + p3;
    return (p1 * min_depth + p2 * min_depth + p3) / error;
}

// ----------------------------------------------------------------------------------------------
// correspondences
// ----------------------------------------------------------------------------------------------
struct Corr {
    int q, m;
    float dist;  // pcl::Correspondence::distance
};

// PCL 1.14 CorrespondenceRejectorTrimmed::getRemainingCorrespondences restated
// (reference .cpp:487-488,508-510,634-635,669-671): overlap stored as float,
// n_keep = floor(ratio_f32 * float(N)); pass-through in original order if n_keep >= N, otherwise
// std::nth_element + resize.  keep_largest selects the comparator direction.
size_t trimmed_count(size_t n, double overlap) {
    float ratio = (float)overlap;
    float prod = ratio * static_cast<float>(n);
    double fl = std::floor((double)prod);
    if (fl < 0.0) fl = 0.0;
    return (size_t)(unsigned int)fl;
}

void trim_correspondences(const std::vector<Corr>& in, double overlap, bool keep_largest, std::vector<Corr>& out) {
    size_t n_keep = trimmed_count(in.size(), overlap);
    if (n_keep < in.size()) {
        out = in;
        if (keep_largest)
            std::nth_element(out.begin(), out.begin() + n_keep, out.end(),
                             [](const Corr& a, const Corr& b) { return a.dist > b.dist; });
        else
            std::nth_element(out.begin(), out.begin() + n_keep, out.end(),
                             [](const Corr& a, const Corr& b) { return a.dist < b.dist; });
        out.resize(n_keep);
    } else {
        out = in;
    }
}

// reference .cpp:379-387 mean of the stored float distances
double mean_stored_distance(const std::vector<Corr>& c) {
    double s = 0.0;
    int n = 0;
    for (const Corr& k : c) {
        s += k.dist;
        n++;
    }
    return s / n;
}
// reference .cpp:390-400
double mean_euclidean_distance(const Cloud& src, const Cloud& tgt, const std::vector<Corr>& c) {
    double s = 0.0;
    int n = 0;
    for (const Corr& k : c) {
        s += norm(src.p(k.q) - tgt.p(k.m));
        n++;
    }
    return s / n;
}

// reference .cpp:402-416
void correspondences_xyz(const Cloud& moving, const KDTree& target_tree, std::vector<Corr>& out) {
    size_t n = moving.size();
    out.resize(n);
#pragma omp parallel for schedule(dynamic, 256)
    for (long i = 0; i < (long)n; i++) {
        Neighbor nb;
        target_tree.knn(&moving.pts[3 * i], 1, &nb);
        out[i] = {(int)i, nb.idx, float(std::sqrt(nb.d2))};
    }
}

inline void se3_query(const M4& X, double q[12]) {  // reference .cpp:450-453 (R column-major, then p)
    q[0] = X(0, 0), q[1] = X(1, 0), q[2] = X(2, 0);
    q[3] = X(0, 1), q[4] = X(1, 1), q[5] = X(2, 1);
    q[6] = X(0, 2), q[7] = X(1, 2), q[8] = X(2, 2);
    q[9] = X(0, 3), q[10] = X(1, 3), q[11] = X(2, 3);
}

// reference .cpp:444-470: 12-D search, stored distance = 3-D distance of the translation columns
void correspondences_se3(const std::vector<M4>& src_se3, const std::vector<M4>& tgt_se3, const KDTree& tree12,
                         std::vector<Corr>& out) {
    size_t n = src_se3.size();
    out.resize(n);
#pragma omp parallel for schedule(dynamic, 64)
    for (long i = 0; i < (long)n; i++) {
        double q[12];
        se3_query(src_se3[i], q);
        Neighbor nb;
        tree12.knn(q, 1, &nb);
        const M4& S = src_se3[i];
        const M4& Tm = tgt_se3[nb.idx];
        V3 d{S(0, 3) - Tm(0, 3), S(1, 3) - Tm(1, 3), S(2, 3) - Tm(2, 3)};
        out[i] = {(int)i, nb.idx, float(norm(d))};
    }
}

// ----------------------------------------------------------------------------------------------
// estimators
// ----------------------------------------------------------------------------------------------
struct NormalEq {
    double JTJ[36];
    double JTr[6];
    void zero() {
        std::memset(JTJ, 0, sizeof(JTJ));
        std::memset(JTr, 0, sizeof(JTr));
    }
    inline void add_row(const double J[6], double r, double w) {  // Open3D ComputeJTJandJTr body
        for (int a = 0; a < 6; a++) {
            for (int b = 0; b < 6; b++) JTJ[6 * a + b] += J[a] * w * J[b];
            JTr[a] += J[a] * w * r;
        }
    }
    void merge(const NormalEq& o) {
        for (int i = 0; i < 36; i++) JTJ[i] += o.JTJ[i];
        for (int i = 0; i < 6; i++) JTr[i] += o.JTr[i];
    }
};

// Open3D TransformVector6dToMatrix4d: R = Rz(x2) Ry(x1) Rx(x0), t = x3..5
M4 vector6_to_matrix(const double x[6]) {
    double cx = std::cos(x[0]), sx = std::sin(x[0]);
    double cy = std::cos(x[1]), sy = std::sin(x[1]);
    double cz = std::cos(x[2]), sz = std::sin(x[2]);
    M3 Rx = m3_identity(), Ry = m3_identity(), Rz = m3_identity();
    Rx(1, 1) = cx, Rx(1, 2) = -sx, Rx(2, 1) = sx, Rx(2, 2) = cx;
    Ry(0, 0) = cy, Ry(0, 2) = sy, Ry(2, 0) = -sy, Ry(2, 2) = cy;
    Rz(0, 0) = cz, Rz(0, 1) = -sz, Rz(1, 0) = sz, Rz(1, 1) = cz;
    M3 R = m3_mul(m3_mul(Rz, Ry), Rx);
    M4 T = m4_identity();
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) T(i, j) = R(i, j);
    T(0, 3) = x[3], T(1, 3) = x[4], T(2, 3) = x[5];
    return T;
}

// Open3D SolveLinearSystemPSD(JTJ, -JTr) -> Eigen LDLT.  Restated as a pivoted symmetric LDL^T
// (diagonal pivoting, as Eigen's LDLT does); returns false on a non-finite solution.
bool solve_ldlt6(const double A_in[36], const double b_in[6], double x[6]) {
    const int n = 6;
    double A[36], b[6];
    std::memcpy(A, A_in, sizeof(A));
    std::memcpy(b, b_in, sizeof(b));
    int perm[6];
    for (int i = 0; i < n; i++) perm[i] = i;
    double L[36] = {0}, D[6] = {0};
    for (int k = 0; k < n; k++) {
        int piv = k;
        double big = std::fabs(A[6 * k + k]);
        for (int i = k + 1; i < n; i++)
            if (std::fabs(A[6 * i + i]) > big) {
                big = std::fabs(A[6 * i + i]);
                piv = i;
            }
        if (piv != k) {  // symmetric row/column swap
            for (int j = 0; j < n; j++) std::swap(A[6 * k + j], A[6 * piv + j]);
            for (int i = 0; i < n; i++) std::swap(A[6 * i + k], A[6 * i + piv]);
            for (int j = 0; j < k; j++) std::swap(L[6 * k + j], L[6 * piv + j]);
            std::swap(perm[k], perm[piv]);
        }
        D[k] = A[6 * k + k];
        L[6 * k + k] = 1.0;
        if (D[k] == 0.0) continue;
        for (int i = k + 1; i < n; i++) L[6 * i + k] = A[6 * i + k] / D[k];
        for (int i = k + 1; i < n; i++)
            for (int j = k + 1; j < n; j++) A[6 * i + j] -= L[6 * i + k] * D[k] * L[6 * j + k];
    }
    double y[6], z[6];
    for (int i = 0; i < n; i++) y[i] = b[perm[i]];
    for (int i = 0; i < n; i++)
        for (int j = 0; j < i; j++) y[i] -= L[6 * i + j] * y[j];
    for (int i = 0; i < n; i++) z[i] = (D[i] != 0.0) ? y[i] / D[i] : 0.0;
    for (int i = n - 1; i >= 0; i--)
        for (int j = i + 1; j < n; j++) z[i] -= L[6 * j + i] * z[j];
    for (int i = 0; i < n; i++) x[perm[i]] = z[i];
    for (int i = 0; i < n; i++)
        if (!std::isfinite(x[i])) return false;
    return true;
}

// Open3D SolveJacobianSystemAndObtainExtrinsicMatrix
M4 solve_normal_equations(const NormalEq& ne) {
    double rhs[6], x[6];
    for (int i = 0; i < 6; i++) rhs[i] = -ne.JTr[i];
    if (!solve_ldlt6(ne.JTJ, rhs, x)) return m4_identity();
    return vector6_to_matrix(x);
}

// Open3D TransformationEstimationPointToPlane::ComputeTransformation (reference .cpp:528,695,1095)
NormalEq reduce_point_to_plane(const double* src, const double* tgt, const double* tgt_n, const Corr* corr, size_t k) {
    NormalEq total;
    total.zero();
#pragma omp parallel
    {
        NormalEq local;
        local.zero();
#pragma omp for schedule(static) nowait
        for (long i = 0; i < (long)k; i++) {
            V3 vs = load3(&src[3 * corr[i].q]);
            V3 vt = load3(&tgt[3 * corr[i].m]);
            V3 nt = load3(&tgt_n[3 * corr[i].m]);
            double r = dot(vs - vt, nt);
            V3 c = cross(vs, nt);
            double J[6] = {c.x, c.y, c.z, nt.x, nt.y, nt.z};
            local.add_row(J, r, 1.0);
        }
#pragma omp critical
        total.merge(local);
    }
    return total;
}

// reference .cpp:57-110 optimize_generalizedICP_manual (== Open3D ForGeneralizedICP when w == 1):
// M = Ct + Cs, W = w * (M^-1)^(1/2), J = W [-[vs]x | I], r = W (vs - vt), three rows each.
NormalEq reduce_gicp(const double* src, const double* src_cov, const double* tgt, const double* tgt_cov,
                     const Corr* corr, const double* weights, size_t k) {
    NormalEq total;
    total.zero();
#pragma omp parallel
    {
        NormalEq local;
        local.zero();
#pragma omp for schedule(static) nowait
        for (long i = 0; i < (long)k; i++) {
            V3 vs = load3(&src[3 * corr[i].q]);
            V3 vt = load3(&tgt[3 * corr[i].m]);
            M3 M;
            for (int e = 0; e < 9; e++) M.a[e] = tgt_cov[9 * (size_t)corr[i].m + e] + src_cov[9 * (size_t)corr[i].q + e];
            // principal square root of M^-1 via the eigen-decomposition of the SPD matrix M
            double ev[3];
            M3 V;
            eig3_sym(M, ev, V);
            M3 W = m3_zero();
            for (int e = 0; e < 3; e++) {
                double s = 1.0 / std::sqrt(ev[e]);
                for (int r = 0; r < 3; r++)
                    for (int c = 0; c < 3; c++) W(r, c) += s * V(r, e) * V(c, e);
            }
            double wgt = weights ? weights[i] : 1.0;
            for (double& v : W.a) v *= wgt;
            M3 negskew = m3_skew(vs);
            for (double& v : negskew.a) v = -v;
            M3 WJ = m3_mul(W, negskew);
            V3 d = vs - vt;
            V3 r = m3_mulv(W, d);
            double rr[3] = {r.x, r.y, r.z};
            for (int row = 0; row < 3; row++) {
                double J[6] = {WJ(row, 0), WJ(row, 1), WJ(row, 2), W(row, 0), W(row, 1), W(row, 2)};
                local.add_row(J, rr[row], 1.0);
            }
        }
#pragma omp critical
        total.merge(local);
    }
    return total;
}

// 3x3 SVD A = U S V^T through the Jacobi eigen-decomposition of A^T A (enough for Kabsch).
void svd3(const M3& A, M3& U, double S[3], M3& V) {
    M3 AtA = m3_mul(m3_transpose(A), A);
    double ev[3];
    M3 Va;
    eig3_sym(AtA, ev, Va);  // ascending
    // descending singular values
    for (int c = 0; c < 3; c++) {
        int src = 2 - c;
        S[c] = std::sqrt(std::max(ev[src], 0.0));
        for (int r = 0; r < 3; r++) V(r, c) = Va(r, src);
    }
    V3 u[3];
    for (int c = 0; c < 3; c++) {
        V3 v{V(0, c), V(1, c), V(2, c)};
        u[c] = m3_mulv(A, v);
    }
    double n0 = norm(u[0]);
    u[0] = n0 > 0 ? (1.0 / n0) * u[0] : V3{1, 0, 0};
    // Gram-Schmidt keeps U orthonormal even when the smaller singular values are (near) zero
    u[1] = u[1] - dot(u[1], u[0]) * u[0];
    double n1 = norm(u[1]);
    if (n1 > 1e-14 * std::max(S[0], 1e-300)) {
        u[1] = (1.0 / n1) * u[1];
    } else {
        V3 t = std::fabs(u[0].x) < 0.9 ? V3{1, 0, 0} : V3{0, 1, 0};
        u[1] = cross(u[0], t);
        u[1] = (1.0 / norm(u[1])) * u[1];
    }
    V3 u2 = u[2] - dot(u[2], u[0]) * u[0];
    u2 = u2 - dot(u2, u[1]) * u[1];
    double n2 = norm(u2);
    if (n2 > 1e-14 * std::max(S[0], 1e-300)) {
        u[2] = (1.0 / n2) * u2;
    } else {
        u[2] = cross(u[0], u[1]);
    }
    for (int c = 0; c < 3; c++) U(0, c) = u[c].x, U(1, c) = u[c].y, U(2, c) = u[c].z;
}

// Open3D TransformationEstimationPointToPoint::ComputeTransformation = Eigen::umeyama(src, tgt, false)
// (reference .cpp:525,692,1092): sigma = (1/K) sum (t - mu_t)(s - mu_s)^T, R = U diag(1,1,d) V^T,
// d = sign(det U det V), t = mu_t - R mu_s.
M4 umeyama_no_scale(const double* src, const double* tgt, const Corr* corr, size_t k) {
    if (k == 0) return m4_identity();
    V3 ms{0, 0, 0}, mt{0, 0, 0};
    for (size_t i = 0; i < k; i++) {
        ms = ms + load3(&src[3 * corr[i].q]);
        mt = mt + load3(&tgt[3 * corr[i].m]);
    }
    ms = (1.0 / (double)k) * ms;
    mt = (1.0 / (double)k) * mt;
    M3 sigma = m3_zero();
    for (size_t i = 0; i < k; i++) {
        V3 a = load3(&tgt[3 * corr[i].m]) - mt;
        V3 b = load3(&src[3 * corr[i].q]) - ms;
        double av[3] = {a.x, a.y, a.z}, bv[3] = {b.x, b.y, b.z};
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) sigma(r, c) += av[r] * bv[c];
    }
    for (double& v : sigma.a) v /= (double)k;
    M3 U, V;
    double S[3];
    svd3(sigma, U, S, V);
    double d = (m3_det(U) * m3_det(V) < 0.0) ? -1.0 : 1.0;
    M3 Sd = m3_identity();
    Sd(2, 2) = d;
    M3 R = m3_mul(m3_mul(U, Sd), m3_transpose(V));
    V3 t = mt - m3_mulv(R, ms);
    M4 T = m4_identity();
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) T(i, j) = R(i, j);
    T(0, 3) = t.x, T(1, 3) = t.y, T(2, 3) = t.z;
    return T;
}

M4 estimate_step(int variant, const Cloud& moving, const Cloud& target, const std::vector<Corr>& kept,
                 const double* weights) {
    if (variant == ORC_PT2PT) return umeyama_no_scale(moving.pts.data(), target.pts.data(), kept.data(), kept.size());
    if (variant == ORC_PT2PL) {
        if (kept.empty() || target.normals.empty()) return m4_identity();
        NormalEq ne = reduce_point_to_plane(moving.pts.data(), target.pts.data(), target.normals.data(), kept.data(),
                                            kept.size());
        return solve_normal_equations(ne);
    }
    if (kept.empty() || target.covs.empty() || moving.covs.empty()) return m4_identity();  // .cpp:63-66
    NormalEq ne = reduce_gicp(moving.pts.data(), moving.covs.data(), target.pts.data(), target.covs.data(), kept.data(),
                              weights, kept.size());
    return solve_normal_equations(ne);
}

void record_trace(orc_trace* tr, int iter_index, const std::vector<Corr>& all, size_t n_kept, const M4& Ti, double mean,
                  bool se3) {
    if (!tr || iter_index >= tr->max_iters) return;
    size_t n = all.size();
    if (tr->corr_idx)
        for (size_t i = 0; i < n; i++) tr->corr_idx[(size_t)iter_index * n + i] = all[i].m;
    if (tr->corr_dist)
        for (size_t i = 0; i < n; i++) tr->corr_dist[(size_t)iter_index * n + i] = all[i].dist;
    if (tr->T_iter) std::memcpy(&tr->T_iter[16 * (size_t)iter_index], Ti.a, sizeof(Ti.a));
    if (tr->mean_dist) tr->mean_dist[iter_index] = mean;
    if (tr->se3_phase) tr->se3_phase[iter_index] = se3 ? 1 : 0;
    if (tr->n_kept) tr->n_kept[iter_index] = (int)n_kept;
    tr->n_iters = iter_index + 1;
}

// ----------------------------------------------------------------------------------------------
// reference .cpp:473-552 run_icp
// ----------------------------------------------------------------------------------------------
int run_icp(Cloud& source_moving, Cloud& target, const orc_params& P, M4& T_total, orc_stats& st, orc_trace* tr) {
    auto t_all = Clock::now();
    KDTree tree_t;
    tree_t.build(target.pts.data(), target.size(), 3);  // .cpp:482
    T_total = m4_identity();
    double mse_prev = 1e7, mse_cur = 1e7, mse_rel = 1e7;
    if (P.variant == ORC_PT2PL) estimate_normals(target, tree_t, P.knn_normals_pt2pl);  // .cpp:493-495
    if (P.variant == ORC_GICP) {                                                        // .cpp:497-500
        KDTree tree_s;
        tree_s.build(source_moving.pts.data(), source_moving.size(), 3);
        init_gicp(source_moving, tree_s, P.knn_normals_gicp, P.gicp_epsilon);
        init_gicp(target, tree_t, P.knn_normals_gicp, P.gicp_epsilon);
    }
    st.time_setup_ms = ms_since(t_all);
    int it = 0;
    std::vector<Corr> all, kept;
    while (true) {
        auto t0 = Clock::now();
        correspondences_xyz(source_moving, tree_t, all);  // .cpp:505
        st.time_corr_ms += ms_since(t0);
        t0 = Clock::now();
        trim_correspondences(all, P.estimated_overlap, P.trim_keep_largest != 0, kept);  // .cpp:508-510
        mse_prev = mse_cur;
        mse_cur = mean_stored_distance(kept);  // .cpp:519-521
        mse_rel = std::fabs(mse_cur - mse_prev);
        M4 Ti = estimate_step(P.variant, source_moving, target, kept, nullptr);  // .cpp:523-535
        record_trace(tr, it, all, kept.size(), Ti, mse_cur, false);
        cloud_transform(source_moving, Ti);  // .cpp:541
        T_total = m4_mul(Ti, T_total);       // .cpp:544
        st.time_opt_ms += ms_since(t0);
        it++;  // .cpp:547
        if (it == P.max_num_iterations || mse_rel < P.mse) break;
    }
    st.num_iterations = it;
    st.num_pure_se3_iterations = -1;
    st.scaling_factor = 1.0;
    st.time_total_ms = ms_since(t_all);
    return 0;
}

// ----------------------------------------------------------------------------------------------
// reference .cpp:555-739 run_se3_icp, :742-959 run_se3_icp_with_cf, :962-1128 run_se3_pure
// ----------------------------------------------------------------------------------------------
int run_se3(Cloud& source, Cloud& source_moving, Cloud& target, const orc_params& P, M4& T_total, orc_stats& st,
            orc_trace* tr) {
    auto t_all = Clock::now();
    const bool with_cf = P.entry == ORC_RUN_SE3_ICP_CF;
    const bool pure = P.entry == ORC_RUN_SE3_PURE;
    const int variant = with_cf ? (int)ORC_GICP : P.variant;
    size_t n = source.size(), m = target.size();

    // .cpp:756-769 confidences from the raw (un-normalised) depth
    std::vector<double> conf_s, conf_t;
    if (with_cf) {
        conf_s.resize(n);
        conf_t.resize(m);
        for (size_t i = 0; i < n; i++) conf_s[i] = lounge_point_confidence(source.p(i));
        for (size_t i = 0; i < m; i++) conf_t[i] = lounge_point_confidence(target.p(i));
    }

    // .cpp:568-582 normalisation
    V3 c_s = cloud_center(source), c_t = cloud_center(target);
    double r_s = largest_distance_from(c_s, source), r_t = largest_distance_from(c_t, target);
    double r_max = std::max(r_s, r_t);
    double s = P.scale_preprocessing * (1.0 / r_max);
    cloud_translate(source, -1.0 * c_s);
    cloud_translate(source_moving, -1.0 * c_s);
    cloud_translate(target, -1.0 * c_t);
    cloud_scale(source, s);
    cloud_scale(source_moving, s);
    cloud_scale(target, s);

    // .cpp:586-591 trees and TOLDI frames
    KDTree tree_s, tree_t;
    tree_s.build(source.pts.data(), n, 3);
    tree_t.build(target.pts.data(), m, 3);
    std::vector<M4> src_se3, tgt_se3;
    if (P.lrf_method == 1) {  // .cpp:593-594 (commented out in the reference)
        shot_all(source, tree_s, P.lrf_radius, src_se3);
        shot_all(target, tree_t, P.lrf_radius, tgt_se3);
    } else {
        toldi_all(source, tree_s, P.number_of_nn_for_LRF, src_se3);
        toldi_all(target, tree_t, P.number_of_nn_for_LRF, tgt_se3);
    }

    // .cpp:597-607 rotation block * alpha, translation column * beta
    auto weight = [&](std::vector<M4>& v) {
#pragma omp parallel for schedule(static)
        for (long i = 0; i < (long)v.size(); i++) {
            for (int r = 0; r < 3; r++) {
                for (int c = 0; c < 3; c++) v[i](r, c) *= P.alpha_rot;
                v[i](r, 3) *= P.beta_transl;
            }
        }
    };
    weight(src_se3);
    weight(tgt_se3);

    // .cpp:610-626 12 x M data matrix and its tree (with_cf takes rows 9-11 from target_.points_, .cpp:834-836)
    std::vector<double> rows_t(12 * m);
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)m; i++) {
        se3_query(tgt_se3[i], &rows_t[12 * i]);
        if (with_cf) {
            rows_t[12 * i + 9] = target.pts[3 * i];
            rows_t[12 * i + 10] = target.pts[3 * i + 1];
            rows_t[12 * i + 11] = target.pts[3 * i + 2];
        }
    }
    KDTree tree12;
    tree12.build(rows_t.data(), m, 12);

    T_total = m4_identity();
    M4 T_prev = m4_identity();
    double mse_prev = 1e7, mse_cur = 1e7, mse_rel = 1e7, T_change = 1e7;
    int it = 0, se3_it = 0;

    // .cpp:642-648 variant set-up
    if (variant == ORC_PT2PL) {
        estimate_normals(target, tree_t, P.knn_normals_pt2pl);
    } else if (variant == ORC_GICP) {
        init_gicp(source_moving, tree_s, P.knn_normals_gicp, P.gicp_epsilon);
        init_gicp(target, tree_t, P.knn_normals_gicp, P.gicp_epsilon);
    }
    st.time_setup_ms = ms_since(t_all);

    bool switch_icp = false;
    std::vector<Corr> all, kept;
    std::vector<double> weights;
    while (true) {
        it++;  // .cpp:656
        auto t0 = Clock::now();
        bool se3_search = pure || !switch_icp;
        if (se3_search) {
            se3_it++;
            correspondences_se3(src_se3, tgt_se3, tree12, all);  // .cpp:661
        } else {
            correspondences_xyz(source_moving, tree_t, all);  // .cpp:665
        }
        st.time_corr_ms += ms_since(t0);
        t0 = Clock::now();
        trim_correspondences(all, P.estimated_overlap, P.trim_keep_largest != 0, kept);  // .cpp:669-671
        mse_prev = mse_cur;
        mse_cur = with_cf ? mean_euclidean_distance(source_moving, target, kept)  // .cpp:897
                          : mean_stored_distance(kept);                            // .cpp:685
        mse_rel = std::fabs(mse_cur - mse_prev);
        const double* wptr = nullptr;
        if (with_cf) {  // .cpp:911-919 — the 0.15 "filter" has no effect on what the solver receives
            weights.resize(kept.size());
            for (size_t ci = 0; ci < kept.size(); ci++) weights[ci] = (conf_s[kept[ci].q] + conf_t[kept[ci].m]) / 2.0;
            wptr = weights.data();
        }
        M4 Ti = estimate_step(variant, source_moving, target, kept, wptr);  // .cpp:691-703 / :921
        record_trace(tr, it - 1, all, kept.size(), Ti, mse_cur, se3_search);
        cloud_transform(source_moving, Ti);  // .cpp:706
        T_prev = T_total;
        T_total = m4_mul(Ti, T_total);  // .cpp:710
        T_change = m4_diff_fro(T_prev, T_total);
#pragma omp parallel for schedule(static)
        for (long k = 0; k < (long)src_se3.size(); k++) src_se3[k] = m4_mul(Ti, src_se3[k]);  // .cpp:713-716
        st.time_opt_ms += ms_since(t0);

        if (pure) {  // .cpp:1118
            if (it == P.max_num_se3_iterations || mse_rel < s * P.mse) break;
        } else if (!switch_icp) {  // .cpp:718-723
            if (it == P.max_num_se3_iterations || T_change < P.mse_switch_error) {
                switch_icp = true;
                st.time_before_pure_icp_ms = ms_since(t_all);
            }
        } else {  // .cpp:724-729
            if (it == P.max_num_iterations || mse_rel < s * P.mse) break;
        }
    }

    // .cpp:735-738 back to the original coordinates and scale
    M3 R;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) R(i, j) = T_total(i, j);
    V3 t_prime{T_total(0, 3), T_total(1, 3), T_total(2, 3)};
    V3 t_og = (1.0 / s) * t_prime - m3_mulv(R, c_s) + c_t;
    T_total(0, 3) = t_og.x, T_total(1, 3) = t_og.y, T_total(2, 3) = t_og.z;

    st.num_iterations = it;
    st.num_pure_se3_iterations = se3_it;
    st.scaling_factor = s;
    st.time_total_ms = ms_since(t_all);
    return 0;
}

void fill_cloud(Cloud& c, const double* xyz, size_t n) { c.pts.assign(xyz, xyz + 3 * n); }

}  // namespace

// ================================================================================================
// C interface
// ================================================================================================
extern "C" {

void orc_default_params(orc_params* p) {  // reference ctor .cpp:334-348
    p->variant = ORC_PT2PL;
    p->entry = ORC_RUN_SE3_ICP;
    p->max_num_iterations = 150;
    p->max_num_se3_iterations = 20;
    p->number_of_nn_for_LRF = 30;
    p->knn_normals_pt2pl = 30;
    p->knn_normals_gicp = 20;
    p->trim_keep_largest = 1;
    p->mse = 0.00001;
    p->mse_switch_error = 0.001;
    p->estimated_overlap = 1.0;
    p->alpha_rot = 3.0;
    p->beta_transl = 1.0;
    p->scale_preprocessing = 3.0;
    p->gicp_epsilon = 1e-3;
    p->lrf_method = 0;
    p->reserved0 = 0;
    p->lrf_radius = 0.8;
}

int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_run(const double* src_xyz, size_t n, const double* tgt_xyz, size_t m, const orc_params* p, double* T_out,
            orc_stats* stats, orc_trace* trace) {
    if (!src_xyz || !tgt_xyz || !p || !T_out || n == 0 || m == 0) return 1;
    if (p->variant < ORC_PT2PT || p->variant > ORC_GICP) return 2;
    Cloud source, moving, target;
    fill_cloud(source, src_xyz, n);
    fill_cloud(moving, src_xyz, n);
    fill_cloud(target, tgt_xyz, m);
    orc_stats st;
    std::memset(&st, 0, sizeof(st));
    M4 T;
    if (trace) trace->n_iters = 0;
    int rc;
    if (p->entry == ORC_RUN_ICP)
        rc = run_icp(moving, target, *p, T, st, trace);
    else
        rc = run_se3(source, moving, target, *p, T, st, trace);
    std::memcpy(T_out, T.a, sizeof(T.a));
    if (stats) *stats = st;
    return rc;
}

int orc_knn_self(const double* xyz, size_t n, int k, int32_t* idx, double* d2) {
    KDTree tree;
    tree.build(xyz, n, 3);
    int kk = (int)std::min<size_t>(k, n);
#pragma omp parallel
    {
        std::vector<Neighbor> nb(k);
#pragma omp for schedule(dynamic, 64)
        for (long i = 0; i < (long)n; i++) {
            int cnt = tree.knn(&xyz[3 * i], k, nb.data());
            for (int j = 0; j < k; j++) {
                idx[(size_t)i * k + j] = j < cnt ? nb[j].idx : -1;
                if (d2) d2[(size_t)i * k + j] = j < cnt ? nb[j].d2 : -1.0;
            }
        }
    }
    return kk;
}

int orc_toldi(const double* xyz, size_t n, int k, double* frames) {
    Cloud c;
    fill_cloud(c, xyz, n);
    KDTree tree;
    tree.build(xyz, n, 3);
    std::vector<M4> f;
    toldi_all(c, tree, k, f);
    for (size_t i = 0; i < n; i++) std::memcpy(&frames[16 * i], f[i].a, sizeof(f[i].a));
    return 0;
}

int orc_shot(const double* xyz, size_t n, double radius, double* frames) {
    Cloud c;
    fill_cloud(c, xyz, n);
    KDTree tree;
    tree.build(xyz, n, 3);
    std::vector<M4> f;
    shot_all(c, tree, radius, f);
    for (size_t i = 0; i < n; i++) std::memcpy(&frames[16 * i], f[i].a, sizeof(f[i].a));
    return 0;
}

int orc_normals(const double* xyz, size_t n, int k, double* normals) {
    Cloud c;
    fill_cloud(c, xyz, n);
    KDTree tree;
    tree.build(xyz, n, 3);
    estimate_normals(c, tree, k);
    std::memcpy(normals, c.normals.data(), sizeof(double) * 3 * n);
    return 0;
}

int orc_gicp_cov(const double* normals, size_t n, double eps, double* cov) {
    std::vector<double> nrm(normals, normals + 3 * n), out;
    gicp_covariances_from_normals(nrm, eps, out);
    std::memcpy(cov, out.data(), sizeof(double) * 9 * n);
    return 0;
}

int orc_se3_rows(const double* frames, size_t n, double alpha, double beta, double* rows) {
    for (size_t i = 0; i < n; i++) {
        M4 X;
        std::memcpy(X.a, &frames[16 * i], sizeof(X.a));
        for (int r = 0; r < 3; r++) {
            for (int c = 0; c < 3; c++) X(r, c) *= alpha;
            X(r, 3) *= beta;
        }
        se3_query(X, &rows[12 * i]);
    }
    return 0;
}

int orc_nn(const double* queries, size_t nq, const double* data, size_t nd, int dim, int32_t* idx, double* d2) {
    KDTree tree;
    tree.build(data, nd, dim);
#pragma omp parallel for schedule(dynamic, 64)
    for (long i = 0; i < (long)nq; i++) {
        Neighbor nb;
        tree.knn(&queries[(size_t)i * dim], 1, &nb);
        idx[i] = nb.idx;
        if (d2) d2[i] = nb.d2;
    }
    return 0;
}

int orc_nn_brute(const double* queries, size_t nq, const double* data, size_t nd, int dim, int32_t* idx, double* d2) {
#pragma omp parallel for schedule(dynamic, 16)
    for (long i = 0; i < (long)nq; i++) {
        const double* q = &queries[(size_t)i * dim];
        double best = std::numeric_limits<double>::infinity();
        int bi = -1;
        for (size_t j = 0; j < nd; j++) {
            const double* p = &data[j * dim];
            double s = 0.0;
            for (int d = 0; d < dim; d++) {
                double df = q[d] - p[d];
                s += df * df;
            }
            if (s < best) {
                best = s;
                bi = (int)j;
            }
        }
        idx[i] = bi;
        if (d2) d2[i] = best;
    }
    return 0;
}

int orc_trim(const float* dist, size_t n, double overlap, int keep_largest, uint8_t* keep) {
    std::vector<Corr> all(n), kept;
    for (size_t i = 0; i < n; i++) all[i] = {(int)i, 0, dist[i]};
    trim_correspondences(all, overlap, keep_largest != 0, kept);
    std::memset(keep, 0, n);
    for (const Corr& c : kept) keep[c.q] = 1;
    return (int)kept.size();
}

static void pack27(const NormalEq& ne, double* out27) {
    int o = 0;
    for (int a = 0; a < 6; a++)
        for (int b = a; b < 6; b++) out27[o++] = ne.JTJ[6 * a + b];
    for (int a = 0; a < 6; a++) out27[o++] = ne.JTr[a];
}

static std::vector<Corr> make_corr(const int32_t* cs, const int32_t* ct, size_t k) {
    std::vector<Corr> c(k);
    for (size_t i = 0; i < k; i++) c[i] = {cs[i], ct[i], 0.0f};
    return c;
}

int orc_reduce_pt2pl(const double* src, const double* tgt, const double* tgt_normals, const int32_t* corr_src,
                     const int32_t* corr_tgt, size_t k, double* out27) {
    auto c = make_corr(corr_src, corr_tgt, k);
    pack27(reduce_point_to_plane(src, tgt, tgt_normals, c.data(), k), out27);
    return 0;
}

int orc_reduce_gicp(const double* src, const double* src_cov, const double* tgt, const double* tgt_cov,
                    const int32_t* corr_src, const int32_t* corr_tgt, const double* weights, size_t k, double* out27) {
    auto c = make_corr(corr_src, corr_tgt, k);
    pack27(reduce_gicp(src, src_cov, tgt, tgt_cov, c.data(), weights, k), out27);
    return 0;
}

int orc_solve6(const double* in27, double* T_out) {
    NormalEq ne;
    ne.zero();
    int o = 0;
    for (int a = 0; a < 6; a++)
        for (int b = a; b < 6; b++) {
            ne.JTJ[6 * a + b] = in27[o];
            ne.JTJ[6 * b + a] = in27[o];
            o++;
        }
    for (int a = 0; a < 6; a++) ne.JTr[a] = in27[o++];
    M4 T = solve_normal_equations(ne);
    std::memcpy(T_out, T.a, sizeof(T.a));
    return 0;
}

int orc_umeyama(const double* src, const double* tgt, const int32_t* corr_src, const int32_t* corr_tgt, size_t k,
                double* T_out) {
    auto c = make_corr(corr_src, corr_tgt, k);
    M4 T = umeyama_no_scale(src, tgt, c.data(), k);
    std::memcpy(T_out, T.a, sizeof(T.a));
    return 0;
}

int orc_eig3(const double* A, double* evals, double* evecs) {
    M3 M, V;
    std::memcpy(M.a, A, sizeof(M.a));
    eig3_sym(M, evals, V);
    std::memcpy(evecs, V.a, sizeof(V.a));
    return 0;
}

}  // extern "C"
