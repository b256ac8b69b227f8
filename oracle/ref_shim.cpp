// TEST INFRASTRUCTURE ONLY.
//
// C entry points around the reference's OWN registration class, for oracle/_ref/libse3icp_reference.so: this file is
// compiled together with /root/reference/src/iterative_SE3_registration.cpp (unmodified, where it lies) against the
// stand-in headers in compat/ plus the third-party restatements in oracle/refdeps/ (see oracle/Makefile).  It lets the
// tests run the reference's control flow — TOLDI frames, 12-D rows, phase switch, stop tests, un-normalisation — and
// compare oracle/se3icp_oracle.cpp (and, on the GPU box, the CUDA path) against it.
#include <cstring>
#include <string>

#include "iterative_SE3_registration.hpp"

// defined (external linkage, not declared in the header) in the reference's src/iterative_SE3_registration.cpp:318-331
void computeAllTOLDISE3FramesOMP(const open3d::geometry::PointCloud& cloud, const open3d::geometry::KDTreeFlann& kdtree_for_LRF,
                                 int knn_pts, std::vector<Eigen::Matrix4d>& result_frames);
// reference .cpp:226-239 (same linkage; its call sites in the class are commented out, .cpp:593-594)
void computeAllSHOTSE3FramesOMP(const open3d::geometry::PointCloud& cloud, const open3d::geometry::KDTreeFlann& kdtree_for_LRF,
                                double radius, std::vector<Eigen::Matrix4d>& result_frames);
// reference .cpp:33-52
void InitializePointCloudForGeneralizedICP_modified(open3d::geometry::PointCloud& pcd, double epsilon);

namespace {
open3d::geometry::PointCloud make_cloud(const double* xyz, size_t n) {
    open3d::geometry::PointCloud pc;
    pc.points_.resize(n);
    for (size_t i = 0; i < n; i++) pc.points_[i] = Eigen::Vector3d(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]);
    return pc;
}
void store_row_major(const Eigen::Matrix4d& T, double* out) {
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) out[4 * r + c] = T(r, c);
}
}  // namespace

extern "C" {

struct ref_params {
    int max_num_iterations, max_num_se3_iterations, number_of_nn_for_LRF, trim_keep_largest;
    double mse, mse_switch_error, estimated_overlap, alpha_rot, beta_transl, scale_preprocessing;
};

// the constructor defaults of the reference class (.cpp:334-348)
void ref_default_params(ref_params* p) {
    IterativeSE3Registration reg;
    p->max_num_iterations = reg.max_num_iterations_;
    p->max_num_se3_iterations = reg.max_num_se3_iterations_;
    p->number_of_nn_for_LRF = reg.number_of_nn_for_LRF_;
    p->trim_keep_largest = 1;
    p->mse = reg.mse_;
    p->mse_switch_error = reg.mse_switch_error_;
    p->estimated_overlap = reg.estimated_overlap_;
    p->alpha_rot = reg.alpha_rot;
    p->beta_transl = reg.beta_transl;
    p->scale_preprocessing = reg.scale_preprocessing;
}

// entry: 0 run_icp, 1 run_se3_icp, 2 run_se3_icp_with_cf, 3 run_se3_pure; variant: "pt2pt" | "pt2pl" | "gicp".
// T_out: 4x4 row-major.  iters_out: {num_iterations_, num_pure_se3_iterations_}.
// corr_out (optional, n ints): target index of every source point after the LAST correspondence pass.
int ref_run(int entry, const char* variant, const double* src, size_t n, const double* tgt, size_t m, const ref_params* p,
            double* T_out, int* iters_out, int* corr_out) {
    IterativeSE3Registration reg;
    reg.setSourceCloud(make_cloud(src, n));
    reg.setTargetCloud(make_cloud(tgt, m));
    reg.max_num_iterations_ = p->max_num_iterations;
    reg.max_num_se3_iterations_ = p->max_num_se3_iterations;
    reg.number_of_nn_for_LRF_ = p->number_of_nn_for_LRF;
    reg.mse_ = p->mse;
    reg.mse_switch_error_ = p->mse_switch_error;
    reg.estimated_overlap_ = p->estimated_overlap;
    reg.alpha_rot = p->alpha_rot;
    reg.beta_transl = p->beta_transl;
    reg.scale_preprocessing = p->scale_preprocessing;
    pcl::registration::trim_keep_largest() = p->trim_keep_largest;
    const std::string v(variant ? variant : "");
    switch (entry) {
        case 0: reg.run_icp(v); break;
        case 1: reg.run_se3_icp(v); break;
        case 2: reg.run_se3_icp_with_cf(); break;
        case 3: reg.run_se3_pure(v); break;
        default: return 1;
    }
    store_row_major(reg.current_estimated_T_, T_out);
    if (iters_out) iters_out[0] = reg.num_iterations_, iters_out[1] = reg.num_pure_se3_iterations_;
    if (corr_out)
        for (size_t i = 0; i < n; i++) corr_out[i] = reg.current_correspondences_set.correspondences_vec[i](1);
    return 0;
}

// TOLDI frames of every point with the reference's own functions (.cpp:241-331); frames: n x 4x4 row-major
int ref_toldi(const double* xyz, size_t n, int knn, double* frames) {
    open3d::geometry::PointCloud pc = make_cloud(xyz, n);
    open3d::geometry::KDTreeFlann tree(pc);
    std::vector<Eigen::Matrix4d> out;
    computeAllTOLDISE3FramesOMP(pc, tree, knn, out);
    for (size_t i = 0; i < n; i++) store_row_major(out[i], frames + 16 * i);
    return 0;
}

// SHOT frames of every point with the reference's own functions (.cpp:121-239); frames: n x 4x4 row-major
int ref_shot(const double* xyz, size_t n, double radius, double* frames) {
    open3d::geometry::PointCloud pc = make_cloud(xyz, n);
    open3d::geometry::KDTreeFlann tree(pc);
    std::vector<Eigen::Matrix4d> out;
    computeAllSHOTSE3FramesOMP(pc, tree, radius, out);
    for (size_t i = 0; i < n; i++) store_row_major(out[i], frames + 16 * i);
    return 0;
}

// GICP covariances exactly as the reference initialises them (.cpp:33-52; normals from 20 neighbours); cov: n x 3x3
int ref_gicp_cov(const double* xyz, size_t n, double epsilon, double* normals, double* cov) {
    open3d::geometry::PointCloud pc = make_cloud(xyz, n);
    InitializePointCloudForGeneralizedICP_modified(pc, epsilon);
    for (size_t i = 0; i < n; i++)
        for (int r = 0; r < 3; r++) {
            if (normals) normals[3 * i + r] = pc.normals_[i][r];
            for (int c = 0; c < 3; c++) cov[9 * i + 3 * r + c] = pc.covariances_[i](r, c);
        }
    return 0;
}

// One 12-D correspondence pass of the reference (.cpp:444-470) over given frames (n x 4x4 / m x 4x4 row-major, already
// weighted); idx_out[n], dist_out[n] (Euclidean distance of the matched positions)
int ref_nn_se3(const double* src_frames, size_t n, const double* tgt_frames, size_t m, int* idx_out, double* dist_out) {
    IterativeSE3Registration reg;
    auto load = [](const double* f, size_t k, std::vector<Eigen::Matrix4d>& v) {
        v.resize(k);
        for (size_t i = 0; i < k; i++)
            for (int r = 0; r < 4; r++)
                for (int c = 0; c < 4; c++) v[i](r, c) = f[16 * i + 4 * r + c];
    };
    load(src_frames, n, reg.source_se3_cloud_);
    load(tgt_frames, m, reg.target_se3_cloud_);
    Eigen::MatrixXd data(12, (long)m);
    for (size_t i = 0; i < m; i++)
        for (int c = 0; c < 4; c++)
            for (int r = 0; r < 3; r++) data(3 * c + r, (long)i) = reg.target_se3_cloud_[i](r, c);  // .cpp:613-624
    reg.raw_flann_kd_tree_target_SE3.SetMatrixData(data);
    reg.current_correspondences_set.correspondences_vec.resize(n);
    reg.current_correspondences_set.distances_vec.resize(n);
    reg.current_correspondences_set_pcl->resize(n);
    reg.update_correspondences_raw_flann_SE3(reg.raw_flann_kd_tree_target_SE3, reg.source_se3_cloud_);
    for (size_t i = 0; i < n; i++) {
        idx_out[i] = reg.current_correspondences_set.correspondences_vec[i](1);
        dist_out[i] = reg.current_correspondences_set.distances_vec[i];
    }
    return 0;
}

}  // extern "C"
