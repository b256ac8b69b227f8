// iterative_SE3_registration.hpp — drop-in declaration of the reference's registration class.
//
// Source-compatible with reference include/iterative_SE3_registration.hpp:20-99: same free
// function, same struct, same class name, same public method signatures and same public data
// members (drivers configure the object by assigning fields and read current_estimated_T_, see
// reference examples/run_registration_method.cpp:35-60, benchmark_kitti.cpp:128-168).  The
// implementation (se3-icp_b200/host/iterative_SE3_registration.cpp) does no numeric work itself:
// every run_*() call goes through the C ABI of include/se3icp.h into CUDA kernels.
//
// The include list is the reference's, so the header builds against real Open3D / PCL / Eigen when
// they are installed, or against the small stand-ins under compat/ when they are not.
#pragma once

#include <iostream>
#include <memory>
#include <string>
#include <vector>

#include <eigen3/unsupported/Eigen/MatrixFunctions>
#include <pcl/registration/correspondence_rejection_trimmed.h>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

#include "open3d/Open3D.h"

#include <chrono>
#include <thread>

// reference hpp:20 / .cpp:112-119
double largestDistanceFromGivenPoint(const Eigen::Vector3d& ref_point, const open3d::geometry::PointCloud& cloud);

// reference hpp:22-25
struct CorrespondencesSet {
    std::vector<Eigen::Vector2i> correspondences_vec;
    std::vector<double> distances_vec;
};

struct se3icp_ctx;  // C-ABI context (include/se3icp.h)

class IterativeSE3Registration {
public:
    IterativeSE3Registration();
    ~IterativeSE3Registration();
    // copyable like the reference class: a copy takes the fields and clouds and gets a GPU context of its own on first use
    IterativeSE3Registration(const IterativeSE3Registration& other);
    IterativeSE3Registration& operator=(const IterativeSE3Registration& other);

    // ================= configuration: plain public fields, assigned by the drivers before run_*() =================
    // (reference hpp:80-95; constructor defaults .cpp:334-348, repeated in se3icp_default_params)
    int max_num_iterations_;      // 150   total iteration budget
    int max_num_se3_iterations_;  // 20    budget of the SE(3) phase
    int number_of_nn_for_LRF_;    // 30    neighbours of the TOLDI frame
    double mse_;                  // 1e-5  stop threshold on the change of the mean correspondence distance
    double mse_switch_error_;     // 1e-3  phase switch on ||T_prev - T||_F
    double estimated_overlap_;    // 1.0   overlap ratio of the trimmed rejector
    double alpha_rot;             // 3.0   rotation weight of the SE(3) metric
    double beta_transl;           // 1.0   translation weight
    double scale_preprocessing;   // 3.0   clouds are scaled to this radius
    double lrf_radius_;           // 0.8   SHOT frame radius (used only after set_use_shot_lrf(true), as in the reference)

    // ================= results of a run_*() call ====================================================================
    Eigen::Matrix4d current_estimated_T_;          // source -> target, original coordinates
    int num_iterations_;                           // iterations executed
    int num_pure_se3_iterations_;                  // ... of which in the SE(3) phase (-1 for run_icp)
    double time_se3_correspondence_search_;        // ms, run_se3_icp_with_cf (hpp:86)
    double time_before_pure_icp_;                  // ms, run_se3_icp_with_cf (hpp:85)
    std::vector<Eigen::Matrix4d> estimated_history_;  // per-iteration increments (run_icp, hpp:63)

    // ================= clouds =========================================================================================
    // file overloads replace the cloud, cloud overloads append to it (reference hpp:31-34, .cpp:350-376)
    void setSourceCloud(const std::string& ply_file);
    void setTargetCloud(const std::string& ply_file);
    void setSourceCloud(const open3d::geometry::PointCloud& points);
    void setTargetCloud(const open3d::geometry::PointCloud& points);
    open3d::geometry::PointCloud source_;
    open3d::geometry::PointCloud source_moving_;
    open3d::geometry::PointCloud target_;

    // ================= registration entry points (reference hpp:46-50) ================================================
    // variant: "pt2pt" | "pt2pl" | "gicp"; anything else prints the reference's message
    void run_se3_icp(const std::string& variant);   // SE(3) phase, then plain ICP   (.cpp:555-739)
    void run_se3_icp_with_cf();                     // GICP with depth confidences   (.cpp:742-959)
    void run_se3_pure(const std::string& variant);  // SE(3) phase only              (.cpp:962-1128)
    void run_icp(const std::string& variant);       // plain ICP                     (.cpp:473-552)

    // ================= single passes and helpers (reference hpp:36-43) ================================================
    // One correspondence pass on the current state (.cpp:402-470).  The kd-tree arguments exist for signature
    // compatibility only: the search structures live on the GPU.
    void update_correspondences_raw_flann_SE3();
    void update_correspondences_raw_flann_SE3(const open3d::geometry::KDTreeFlann& unused_tree,
                                              const std::vector<Eigen::Matrix4d>& se3_cloud);
    void update_correspondences_kd_tree_XYZ(const open3d::geometry::KDTreeFlann& unused_tree);
    // mean of the stored correspondence distances / of recomputed Euclidean distances (.cpp:379-400)
    double estimate_current_mse(const pcl::Correspondences correspondences);
    double estimate_current_mse_compute_euclidean(const open3d::geometry::PointCloud& src, const open3d::geometry::PointCloud& tgt,
                                                  const pcl::Correspondences correspondences);

    // ================= state the reference keeps in public members ====================================================
    // Filled after a run only when set_mirror_state(true): no reference driver reads them (hpp:59-60,74-75).  With the
    // mirror on, run_*() also leaves source_ / target_ centred and scaled and source_moving_ at the estimate, as the
    // reference does (.cpp:568-582,706), so a second run_*() on the same object starts from the same state as upstream.
    std::vector<Eigen::Matrix4d> source_se3_cloud_;
    std::vector<Eigen::Matrix4d> target_se3_cloud_;
    CorrespondencesSet current_correspondences_set;
    pcl::CorrespondencesPtr current_correspondences_set_pcl;
    // Present so that code naming them still compiles; the GPU path does not use them (hpp:66-68,76-78).
    open3d::geometry::KDTreeFlann kd_tree_source_XYZ;
    open3d::geometry::KDTreeFlann kd_tree_target_XYZ;
    open3d::geometry::KDTreeFlann raw_flann_kd_tree_target_SE3;
    open3d::pipelines::registration::TransformationEstimationPointToPoint o3d_estimator;
    open3d::pipelines::registration::TransformationEstimationPointToPlane o3d_estimator_po2pl;
    open3d::pipelines::registration::TransformationEstimationForGeneralizedICP o3d_estimator_generalized;

    // ================= extensions (not in the reference) ==============================================================
    void set_mirror_state(bool on) { mirror_state_ = on; }            // reproduce the reference's post-run member state
    void set_trim_keep_largest(bool on) { trim_keep_largest_ = on; }  // false: keep the smallest distances instead of PCL's
                                                                      // `distance >` comparator (include/se3icp.h)
    void set_device(int device);                                      // CUDA ordinal (default: SE3ICP_DEVICE or 0)
    // true: lift points with the SHOT frame (radius lrf_radius_, reference .cpp:121-239) instead of the TOLDI frame —
    // what un-commenting the reference's calls at .cpp:593-594 / :812-813 would do (LRF ablations)
    void set_use_shot_lrf(bool on) { use_shot_lrf_ = on; }

private:
    se3icp_ctx* context();
    void run_entry(int entry, const std::string& variant_name);
    void store_correspondences(const std::vector<int>& idx, const std::vector<double>& dist);

    se3icp_ctx* ctx_ = nullptr;
    int device_ = -1;
    bool mirror_state_ = false;
    bool trim_keep_largest_ = true;
    bool use_shot_lrf_ = false;
};
