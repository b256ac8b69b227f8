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
    IterativeSE3Registration(const IterativeSE3Registration&) = delete;
    IterativeSE3Registration& operator=(const IterativeSE3Registration&) = delete;

    // ---- input (reference hpp:31-34, .cpp:350-376): file overloads replace, cloud overloads append
    void setSourceCloud(const std::string& filename);
    void setTargetCloud(const std::string& filename);
    void setSourceCloud(const open3d::geometry::PointCloud& cloud);
    void setTargetCloud(const open3d::geometry::PointCloud& cloud);

    // ---- single correspondence passes (reference hpp:36-38, .cpp:402-470); the tree arguments are
    //      accepted for signature compatibility, the search structures live on the GPU
    void update_correspondences_kd_tree_XYZ(const open3d::geometry::KDTreeFlann& target_kd_tree);
    void update_correspondences_raw_flann_SE3();
    void update_correspondences_raw_flann_SE3(const open3d::geometry::KDTreeFlann& se3_tree,
                                              const std::vector<Eigen::Matrix4d>& cloud_vector);

    // ---- mean correspondence distance (reference hpp:40-43, .cpp:379-400)
    double estimate_current_mse(const pcl::Correspondences pcl_corrs);
    double estimate_current_mse_compute_euclidean(const open3d::geometry::PointCloud& cloud_src,
                                                  const open3d::geometry::PointCloud& cloud_tgt,
                                                  const pcl::Correspondences pcl_corrs);

    // ---- registration entry points (reference hpp:46-50); variant_name: "pt2pt" | "pt2pl" | "gicp"
    void run_icp(const std::string& variant_name);       // reference .cpp:473-552
    void run_se3_icp(const std::string& variant_name);   // reference .cpp:555-739
    void run_se3_icp_with_cf();                          // reference .cpp:742-959
    void run_se3_pure(const std::string& variant_name);  // reference .cpp:962-1128

    // ---- extensions (not in the reference) ------------------------------------------------------------
    // Copies the device-side state the reference keeps in public members (SE(3) clouds, correspondence
    // sets) back to the host after a run.  Off by default: no reference driver reads them.
    void set_mirror_state(bool on) { mirror_state_ = on; }
    // 0 = keep the smallest distances in the trimmed rejector (documented PCL intent), 1 = keep the largest
    void set_trim_keep_largest(bool on) { trim_keep_largest_ = on; }
    // CUDA device ordinal used by this object (default: environment SE3ICP_DEVICE or 0)
    void set_device(int device);

    // ---- public state, as in the reference (hpp:53-98) ------------------------------------------------
    open3d::geometry::PointCloud source_;
    open3d::geometry::PointCloud source_moving_;
    open3d::geometry::PointCloud target_;

    std::vector<Eigen::Matrix4d> source_se3_cloud_;
    std::vector<Eigen::Matrix4d> target_se3_cloud_;

    std::vector<Eigen::Matrix4d> estimated_history_;

    open3d::geometry::KDTreeFlann kd_tree_target_XYZ;
    open3d::geometry::KDTreeFlann kd_tree_source_XYZ;
    open3d::geometry::KDTreeFlann raw_flann_kd_tree_target_SE3;

    CorrespondencesSet current_correspondences_set;
    pcl::CorrespondencesPtr current_correspondences_set_pcl;
    open3d::pipelines::registration::TransformationEstimationPointToPoint o3d_estimator;
    open3d::pipelines::registration::TransformationEstimationPointToPlane o3d_estimator_po2pl;
    open3d::pipelines::registration::TransformationEstimationForGeneralizedICP o3d_estimator_generalized;

    int number_of_nn_for_LRF_;
    double mse_;
    double estimated_overlap_;
    double lrf_radius_;
    double mse_switch_error_;
    double time_before_pure_icp_;
    double time_se3_correspondence_search_;

    double alpha_rot;
    double beta_transl;
    double scale_preprocessing;

    int num_iterations_;
    int max_num_iterations_;
    int max_num_se3_iterations_;
    int num_pure_se3_iterations_;

    // result of a run_*() call
    Eigen::Matrix4d current_estimated_T_;

private:
    se3icp_ctx* context();
    void run_entry(int entry, const std::string& variant_name);
    void store_correspondences(const std::vector<int>& idx, const std::vector<double>& dist);

    se3icp_ctx* ctx_ = nullptr;
    int device_ = -1;
    bool mirror_state_ = false;
    bool trim_keep_largest_ = false;
};
