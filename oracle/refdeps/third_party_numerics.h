// TEST INFRASTRUCTURE ONLY — included from compat/open3d/Open3D.h only under -DSE3ICP_REFERENCE_BUILD, i.e. only when
// oracle/Makefile compiles the reference's own src/iterative_SE3_registration.cpp into oracle/_ref/.  The product
// (libse3icp_cuda.so, libiterative_SE3_registration.so) never sees this file.
//
// CPU restatements of the Open3D 0.19.0 (@1868f4332, README.md:24,37 of the reference) entry points the reference
// source calls and that are not vendored in /root/reference.  Restated from the published algorithms:
//   PointCloud::EstimateNormals(KDTreeSearchParamKNN)        reference call sites .cpp:43,494,643
//   TransformationEstimationPointToPoint::ComputeTransformation  (= Eigen::umeyama without scaling)  .cpp:525,692
//   TransformationEstimationPointToPlane::ComputeTransformation  .cpp:528,695
//   TransformationEstimationForGeneralizedICP::ComputeTransformation  .cpp:531,698 (the reference carries a weighted
//       copy of this one at .cpp:57-110, which pins the formula)
//   utility::ComputeJTJandJTr, utility::SolveJacobianSystemAndObtainExtrinsicMatrix  .cpp:100-107
// Written independently of oracle/se3icp_oracle.cpp (different SVD and LDLT formulations, serial accumulation) so that
// agreement between the two is a check and not a tautology.
#pragma once

#include <tuple>

namespace open3d {
namespace utility {

// Sum over rows of J^T w J, J^T w r and w r^2; `f(i, J_r, r, w)` fills one or more rows (Open3D's multi-row overload).
// Open3D accumulates per OpenMP thread and merges under a critical section (order not reproducible); this is serial.
template <typename MatType, typename VecType, typename F>
std::tuple<MatType, VecType, double> ComputeJTJandJTr(F f, int iteration_num, bool /*verbose*/ = true) {
    MatType JTJ;
    VecType JTr;
    double r2_sum = 0.0;
    std::vector<VecType, std::allocator<VecType>> J_r;
    std::vector<double> r, w;
    for (int i = 0; i < iteration_num; i++) {
        f(i, J_r, r, w);
        for (size_t j = 0; j < r.size(); j++) {
            JTJ += J_r[j] * w[j] * J_r[j].transpose();
            JTr += J_r[j] * w[j] * r[j];
            r2_sum += r[j] * r[j] * w[j];
        }
    }
    return std::make_tuple(JTJ, JTr, r2_sum);
}

// x = A^-1 b through a pivoted LDL^T (what Eigen's A.ldlt().solve(b) does: symmetric pivoting on the largest
// remaining diagonal entry, no PSD check in Open3D's default arguments).
inline Eigen::Vector6d SolveLdlt6(const Eigen::Matrix6d& A_in, const Eigen::Vector6d& b) {
    const int n = 6;
    Eigen::Matrix6d A = A_in;
    int perm[n];
    for (int i = 0; i < n; i++) perm[i] = i;
    for (int k = 0; k < n; k++) {
        int p = k;
        for (int i = k + 1; i < n; i++)
            if (std::fabs(A(i, i)) > std::fabs(A(p, p))) p = i;
        if (p != k) {  // symmetric row/column swap
            for (int j = 0; j < n; j++) std::swap(A(k, j), A(p, j));
            for (int i = 0; i < n; i++) std::swap(A(i, k), A(i, p));
            std::swap(perm[k], perm[p]);
        }
        double d = A(k, k);
        if (d == 0.0) continue;
        for (int i = k + 1; i < n; i++) A(i, k) /= d;  // column k of L
        for (int j = k + 1; j < n; j++)
            for (int i = j; i < n; i++) {
                A(i, j) -= A(i, k) * d * A(j, k);
                A(j, i) = A(i, j);
            }
    }
    double y[n];
    for (int i = 0; i < n; i++) y[i] = b[perm[i]];
    for (int i = 0; i < n; i++)
        for (int j = 0; j < i; j++) y[i] -= A(i, j) * y[j];
    for (int i = 0; i < n; i++) y[i] = A(i, i) != 0.0 ? y[i] / A(i, i) : 0.0;
    for (int i = n - 1; i >= 0; i--)
        for (int j = i + 1; j < n; j++) y[i] -= A(j, i) * y[j];
    Eigen::Vector6d x;
    for (int i = 0; i < n; i++) x[perm[i]] = y[i];
    return x;
}

// Open3D TransformVector6dToMatrix4d: R = Rz(x2) Ry(x1) Rx(x0), t = x[3..5]
inline Eigen::Matrix4d TransformVector6dToMatrix4d(const Eigen::Vector6d& x) {
    Eigen::Matrix4d out = Eigen::Matrix4d::Identity();
    Eigen::Matrix3d R = (Eigen::AngleAxisd(x[2], Eigen::Vector3d::UnitZ()) * Eigen::AngleAxisd(x[1], Eigen::Vector3d::UnitY()) *
                         Eigen::AngleAxisd(x[0], Eigen::Vector3d::UnitX()))
                                .matrix();
    out.block<3, 3>(0, 0) = R;
    out.block<3, 1>(0, 3) = Eigen::Vector3d(x[3], x[4], x[5]);
    return out;
}

// Open3D: solve JTJ x = -JTr (SolveLinearSystemPSD with its default arguments always reports success)
inline std::tuple<bool, Eigen::Matrix4d> SolveJacobianSystemAndObtainExtrinsicMatrix(const Eigen::Matrix6d& JTJ,
                                                                                   const Eigen::Vector6d& JTr) {
    Eigen::Vector6d x = SolveLdlt6(JTJ, -JTr);
    return std::make_tuple(true, TransformVector6dToMatrix4d(x));
}

}  // namespace utility

namespace geometry {

// kNN (query point included) -> covariance from cumulants E[xx^T] - E[x]E[x]^T -> eigenvector of the smallest
// eigenvalue; no orientation step when the cloud had no normals; (0,0,1) for degenerate neighbourhoods.
// Open3D's fast path uses an analytic 3x3 eigen-solver; here the Jacobi solver of the Eigen stand-in (same vector up
// to sign and rounding).
inline void PointCloud::EstimateNormals(const KDTreeSearchParam& search_param, bool /*fast_normal_computation*/) {
    int knn = 30;
    if (auto* p = dynamic_cast<const KDTreeSearchParamKNN*>(&search_param)) knn = p->knn_;
    const bool had_normals = HasNormals();
    if (!had_normals) normals_.assign(points_.size(), Eigen::Vector3d(0, 0, 1));
    KDTreeFlann tree(*this);
#pragma omp parallel for schedule(static)
    for (int i = 0; i < (int)points_.size(); i++) {
        std::vector<int> idx;
        std::vector<double> d2;
        Eigen::Vector3d normal(0, 0, 1);
        if (tree.SearchKNN(points_[i], knn, idx, d2) >= 3) {
            double c[9] = {0};
            for (int j : idx) {
                const Eigen::Vector3d& p = points_[j];
                c[0] += p[0], c[1] += p[1], c[2] += p[2];
                c[3] += p[0] * p[0], c[4] += p[0] * p[1], c[5] += p[0] * p[2];
                c[6] += p[1] * p[1], c[7] += p[1] * p[2], c[8] += p[2] * p[2];
            }
            for (double& v : c) v /= (double)idx.size();
            Eigen::Matrix3d cov;
            cov(0, 0) = c[3] - c[0] * c[0], cov(1, 1) = c[6] - c[1] * c[1], cov(2, 2) = c[8] - c[2] * c[2];
            cov(0, 1) = cov(1, 0) = c[4] - c[0] * c[1];
            cov(0, 2) = cov(2, 0) = c[5] - c[0] * c[2];
            cov(1, 2) = cov(2, 1) = c[7] - c[1] * c[2];
            Eigen::SelfAdjointEigenSolver<Eigen::Matrix3d> es;
            es.compute(cov);
            Eigen::Matrix3d V = es.eigenvectors();
            normal = V.col(0);
            if (normal.norm() == 0.0) normal = Eigen::Vector3d(0, 0, 1);
        }
        if (had_normals && normal.dot(normals_[i]) < 0.0) normal = -normal;
        normals_[i] = normal;
    }
}

}  // namespace geometry

namespace pipelines {
namespace registration {

namespace refdeps_detail {
// A = U diag(s) V^T for a 3x3 matrix, singular values descending: eigen-decomposition of A^T A for V, columns of U
// from A v / s (Gram-Schmidt, last column by cross product when its singular value vanishes).
inline void Svd3(const Eigen::Matrix3d& A, Eigen::Matrix3d& U, Eigen::Vector3d& s, Eigen::Matrix3d& V) {
    Eigen::SelfAdjointEigenSolver<Eigen::Matrix3d> es;
    es.compute(A.transpose() * A);
    Eigen::Vector3d w = es.eigenvalues();
    Eigen::Matrix3d E = es.eigenvectors();
    for (int k = 0; k < 3; k++) {  // ascending -> descending
        s[k] = std::sqrt(std::max(0.0, w[2 - k]));
        for (int i = 0; i < 3; i++) V(i, k) = E(i, 2 - k);
    }
    Eigen::Vector3d u[3];
    for (int k = 0; k < 3; k++) {
        u[k] = A * V.col(k);
        for (int j = 0; j < k; j++) u[k] -= u[j] * u[j].dot(u[k]);
        double nrm = u[k].norm();
        if (nrm > 1e-13 * (s[0] > 0 ? s[0] : 1.0)) {
            u[k] = u[k] / nrm;
        } else if (k == 2) {
            u[2] = u[0].cross(u[1]);
        } else {  // rank < 2: any unit vector orthogonal to the previous ones
            Eigen::Vector3d e = std::fabs(k ? u[0][0] : 0.0) < 0.9 ? Eigen::Vector3d(1, 0, 0) : Eigen::Vector3d(0, 1, 0);
            for (int j = 0; j < k; j++) e -= u[j] * u[j].dot(e);
            u[k] = e.normalized();
        }
    }
    for (int k = 0; k < 3; k++)
        for (int i = 0; i < 3; i++) U(i, k) = u[k][i];
}
}  // namespace refdeps_detail

// Eigen::umeyama(source, target, with_scaling = false) over the correspondences
inline Eigen::Matrix4d TransformationEstimationPointToPoint::ComputeTransformation(const geometry::PointCloud& source,
                                                                                   const geometry::PointCloud& target,
                                                                                   const CorrespondenceSet& corres) const {
    if (corres.empty()) return Eigen::Matrix4d::Identity();
    const double n = (double)corres.size();
    Eigen::Vector3d ms, mt;
    for (const auto& c : corres) ms += source.points_[c[0]], mt += target.points_[c[1]];
    ms = ms / n, mt = mt / n;
    Eigen::Matrix3d sigma;
    for (const auto& c : corres) sigma += (target.points_[c[1]] - mt) * (source.points_[c[0]] - ms).transpose();
    sigma = sigma / n;
    Eigen::Matrix3d U, V;
    Eigen::Vector3d sv;
    refdeps_detail::Svd3(sigma, U, sv, V);
    Eigen::Vector3d S(1, 1, 1);
    if (U.determinant() * V.determinant() < 0) S[2] = -1;
    Eigen::Matrix3d R = U * S.asDiagonal() * V.transpose();
    Eigen::Matrix4d T = Eigen::Matrix4d::Identity();
    T.block<3, 3>(0, 0) = R;
    T.block<3, 1>(0, 3) = mt - R * ms;
    return T;
}

inline Eigen::Matrix4d TransformationEstimationPointToPlane::ComputeTransformation(const geometry::PointCloud& source,
                                                                                   const geometry::PointCloud& target,
                                                                                   const CorrespondenceSet& corres) const {
    if (corres.empty() || !target.HasNormals()) return Eigen::Matrix4d::Identity();
    auto row = [&](int i, std::vector<Eigen::Vector6d>& J_r, std::vector<double>& r, std::vector<double>& w) {
        const Eigen::Vector3d& vs = source.points_[corres[i][0]];
        const Eigen::Vector3d& vt = target.points_[corres[i][1]];
        const Eigen::Vector3d& nt = target.normals_[corres[i][1]];
        J_r.resize(1), r.resize(1), w.resize(1);
        r[0] = (vs - vt).dot(nt);
        w[0] = 1.0;  // L2 loss
        Eigen::Vector3d c = vs.cross(nt);
        J_r[0] = Eigen::Vector6d();
        for (int k = 0; k < 3; k++) J_r[0][k] = c[k], J_r[0][3 + k] = nt[k];
    };
    Eigen::Matrix6d JTJ;
    Eigen::Vector6d JTr;
    double r2;
    std::tie(JTJ, JTr, r2) = utility::ComputeJTJandJTr<Eigen::Matrix6d, Eigen::Vector6d>(row, (int)corres.size());
    bool ok;
    Eigen::Matrix4d extrinsic;
    std::tie(ok, extrinsic) = utility::SolveJacobianSystemAndObtainExtrinsicMatrix(JTJ, JTr);
    return ok ? extrinsic : Eigen::Matrix4d::Identity();
}

// Same normal equations as the reference's weighted copy (.cpp:57-110) with all weights 1
inline Eigen::Matrix4d TransformationEstimationForGeneralizedICP::ComputeTransformation(const geometry::PointCloud& source,
                                                                                        const geometry::PointCloud& target,
                                                                                        const CorrespondenceSet& corres) const {
    if (corres.empty() || !target.HasCovariances() || !source.HasCovariances()) return Eigen::Matrix4d::Identity();
    auto rows = [&](int i, std::vector<Eigen::Vector6d>& J_r, std::vector<double>& r, std::vector<double>& w) {
        const Eigen::Vector3d& vs = source.points_[corres[i][0]];
        const Eigen::Vector3d& vt = target.points_[corres[i][1]];
        const Eigen::Vector3d d = vs - vt;
        const Eigen::Matrix3d M = target.covariances_[corres[i][1]] + source.covariances_[corres[i][0]];
        const Eigen::Matrix3d W = M.inverse().sqrt();
        Eigen::Matrix<double, 3, 6> J;
        J.block<3, 3>(0, 0) = -utility::SkewMatrix(vs);
        J.block<3, 3>(0, 3) = Eigen::Matrix3d::Identity();
        J = W * J;
        J_r.resize(3), r.resize(3), w.resize(3);
        for (int k = 0; k < 3; k++) {
            r[k] = W.row(k).dot(d);
            w[k] = 1.0;
            J_r[k] = J.row(k);
        }
    };
    Eigen::Matrix6d JTJ;
    Eigen::Vector6d JTr;
    double r2;
    std::tie(JTJ, JTr, r2) = utility::ComputeJTJandJTr<Eigen::Matrix6d, Eigen::Vector6d>(rows, (int)corres.size());
    bool ok;
    Eigen::Matrix4d extrinsic;
    std::tie(ok, extrinsic) = utility::SolveJacobianSystemAndObtainExtrinsicMatrix(JTJ, JTr);
    return ok ? extrinsic : Eigen::Matrix4d::Identity();
}

}  // namespace registration
}  // namespace pipelines
}  // namespace open3d
