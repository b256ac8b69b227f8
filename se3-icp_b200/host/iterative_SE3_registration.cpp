// Host side of the drop-in class (include/iterative_SE3_registration.hpp): marshals the public
// fields into the POD parameter block of the C ABI (include/se3icp.h), calls libse3icp_cuda.so and
// copies the result back into the same public members the reference fills.  No numeric work on the
// CPU and no fallback: if the CUDA library cannot create a context the run_*() call reports the
// error on stderr and leaves current_estimated_T_ untouched, like the reference does for bad input.
#include "iterative_SE3_registration.hpp"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <utility>

#include "se3icp.h"

namespace {

int variant_from_name(const std::string& v) {
    if (v == "pt2pt") return SE3ICP_PT2PT;
    if (v == "pt2pl") return SE3ICP_PT2PL;
    if (v == "gicp") return SE3ICP_GICP;
    return -1;
}

// std::vector<Eigen::Vector3d> is a dense array of xyz doubles (no padding) in Eigen and in compat/
const double* xyz_ptr(const open3d::geometry::PointCloud& c) {
    static_assert(sizeof(Eigen::Vector3d) == 3 * sizeof(double), "Vector3d must be three packed doubles");
    return c.points_.empty() ? nullptr : reinterpret_cast<const double*>(c.points_.data());
}

void report(const char* where, int rc) {
    std::cerr << "[se3icp] " << where << " failed with status " << rc << ": " << se3icp_last_error() << std::endl;
}

// The reference's drivers build a fresh IterativeSE3Registration per pair (benchmark_kitti.cpp:128).  Creating a
// CUDA context object (stream, pinned buffers, device allocations) costs ~40 ms, a registration of a small cloud
// ~3 ms, so contexts are recycled through a process-wide pool: a destroyed object parks its context (with its
// grown device buffers) and the next object on the same device picks it up.  At most SE3ICP_POOL_CONTEXTS (default 4)
// contexts are parked; the pool never calls CUDA from its static destructor (the runtime may already be unloading at
// process exit, and the driver reclaims everything anyway).
class ContextPool {
public:
    se3icp_ctx* acquire(int device) {
        {
            std::lock_guard<std::mutex> lock(mu_);
            for (size_t i = 0; i < free_.size(); i++)
                if (free_[i].first == device) {
                    se3icp_ctx* c = free_[i].second;
                    free_.erase(free_.begin() + (long)i);
                    return c;
                }
        }
        se3icp_ctx* c = nullptr;
        int rc = se3icp_create(device, nullptr, &c);
        if (rc != SE3ICP_OK) {
            report("se3icp_create", rc);
            return nullptr;
        }
        return c;
    }
    void release(int device, se3icp_ctx* c) {
        if (!c) return;
        std::lock_guard<std::mutex> lock(mu_);
        if (free_.size() < capacity_)
            free_.emplace_back(device, c);
        else
            se3icp_destroy(c);
    }
    ContextPool() {
        const char* e = std::getenv("SE3ICP_POOL_CONTEXTS");
        if (e && *e) capacity_ = (size_t)std::max(0, std::atoi(e));
    }
    ~ContextPool() = default;  // parked contexts are left to process teardown on purpose (see above)

private:
    size_t capacity_ = 4;
    std::mutex mu_;
    std::vector<std::pair<int, se3icp_ctx*>> free_;
};

ContextPool& pool() {
    static ContextPool p;
    return p;
}

Eigen::Matrix4d from_row_major(const double* T) {
    Eigen::Matrix4d M;
    for (int r = 0; r < 4; r++)
        for (int c = 0; c < 4; c++) M(r, c) = T[4 * r + c];
    return M;
}

}  // namespace

double largestDistanceFromGivenPoint(const Eigen::Vector3d& ref_point, const open3d::geometry::PointCloud& cloud) {
    double largest = -1.0;
    for (const auto& p : cloud.points_) largest = std::max(largest, (p - ref_point).norm());
    return largest;
}

IterativeSE3Registration::IterativeSE3Registration()  // defaults of the reference constructor, .cpp:334-348
    : max_num_iterations_(150),
      max_num_se3_iterations_(20),
      number_of_nn_for_LRF_(30),
      mse_(0.00001),
      mse_switch_error_(0.001),
      estimated_overlap_(1.0),
      alpha_rot(3.0),
      beta_transl(1.0),
      scale_preprocessing(3.0),
      lrf_radius_(0.8),
      num_iterations_(0),
      num_pure_se3_iterations_(-1),
      time_se3_correspondence_search_(0.0),
      time_before_pure_icp_(0.0),
      current_correspondences_set_pcl(new pcl::Correspondences) {
    current_estimated_T_.setIdentity();
    const char* dev = std::getenv("SE3ICP_DEVICE");
    device_ = dev ? std::atoi(dev) : 0;
}

IterativeSE3Registration::~IterativeSE3Registration() { pool().release(device_, ctx_); }

IterativeSE3Registration::IterativeSE3Registration(const IterativeSE3Registration& o) : IterativeSE3Registration() { *this = o; }

IterativeSE3Registration& IterativeSE3Registration::operator=(const IterativeSE3Registration& o) {
    if (this == &o) return *this;
    max_num_iterations_ = o.max_num_iterations_;
    max_num_se3_iterations_ = o.max_num_se3_iterations_;
    number_of_nn_for_LRF_ = o.number_of_nn_for_LRF_;
    mse_ = o.mse_;
    mse_switch_error_ = o.mse_switch_error_;
    estimated_overlap_ = o.estimated_overlap_;
    alpha_rot = o.alpha_rot;
    beta_transl = o.beta_transl;
    scale_preprocessing = o.scale_preprocessing;
    lrf_radius_ = o.lrf_radius_;
    current_estimated_T_ = o.current_estimated_T_;
    num_iterations_ = o.num_iterations_;
    num_pure_se3_iterations_ = o.num_pure_se3_iterations_;
    time_se3_correspondence_search_ = o.time_se3_correspondence_search_;
    time_before_pure_icp_ = o.time_before_pure_icp_;
    estimated_history_ = o.estimated_history_;
    source_ = o.source_;
    source_moving_ = o.source_moving_;
    target_ = o.target_;
    source_se3_cloud_ = o.source_se3_cloud_;
    target_se3_cloud_ = o.target_se3_cloud_;
    current_correspondences_set = o.current_correspondences_set;
    current_correspondences_set_pcl.reset(new pcl::Correspondences(*o.current_correspondences_set_pcl));
    mirror_state_ = o.mirror_state_;
    trim_keep_largest_ = o.trim_keep_largest_;
    use_shot_lrf_ = o.use_shot_lrf_;
    set_device(o.device_);  // the GPU context is not shared: this object acquires its own on first use
    return *this;
}

void IterativeSE3Registration::set_device(int device) {
    if (ctx_ && device != device_) {
        pool().release(device_, ctx_);
        ctx_ = nullptr;
    }
    device_ = device;
}

se3icp_ctx* IterativeSE3Registration::context() {
    if (!ctx_) ctx_ = pool().acquire(device_);
    return ctx_;
}

void IterativeSE3Registration::setSourceCloud(const std::string& filename) {
    open3d::io::ReadPointCloud(filename, source_);
    open3d::io::ReadPointCloud(filename, source_moving_);
    current_correspondences_set.correspondences_vec.resize(source_.points_.size());
    current_correspondences_set.distances_vec.resize(source_.points_.size());
    current_correspondences_set_pcl->resize(source_.points_.size());
}

void IterativeSE3Registration::setSourceCloud(const open3d::geometry::PointCloud& cloud) {
    source_.points_.insert(source_.points_.end(), cloud.points_.begin(), cloud.points_.end());
    source_moving_.points_.insert(source_moving_.points_.end(), cloud.points_.begin(), cloud.points_.end());
    current_correspondences_set.correspondences_vec.resize(cloud.points_.size());
    current_correspondences_set.distances_vec.resize(cloud.points_.size());
    current_correspondences_set_pcl->resize(source_.points_.size());
}

void IterativeSE3Registration::setTargetCloud(const std::string& filename) { open3d::io::ReadPointCloud(filename, target_); }

void IterativeSE3Registration::setTargetCloud(const open3d::geometry::PointCloud& cloud) {
    target_.points_.insert(target_.points_.end(), cloud.points_.begin(), cloud.points_.end());
}

double IterativeSE3Registration::estimate_current_mse(const pcl::Correspondences pcl_corrs) {
    double sum = 0.0;
    int n = 0;
    for (const auto& c : pcl_corrs) {
        sum += c.distance;
        n++;
    }
    return sum / n;
}

double IterativeSE3Registration::estimate_current_mse_compute_euclidean(const open3d::geometry::PointCloud& cloud_src,
                                                                        const open3d::geometry::PointCloud& cloud_tgt,
                                                                        const pcl::Correspondences pcl_corrs) {
    double sum = 0.0;
    int n = 0;
    for (const auto& c : pcl_corrs) {
        sum += (cloud_src.points_[c.index_query] - cloud_tgt.points_[c.index_match]).norm();
        n++;
    }
    return sum / n;
}

void IterativeSE3Registration::store_correspondences(const std::vector<int>& idx, const std::vector<double>& dist) {
    size_t n = idx.size();
    current_correspondences_set.correspondences_vec.resize(n);
    current_correspondences_set.distances_vec.resize(n);
    current_correspondences_set_pcl->resize(n);
    for (size_t i = 0; i < n; i++) {
        current_correspondences_set.correspondences_vec[i] = Eigen::Vector2i((int)i, idx[i]);
        current_correspondences_set.distances_vec[i] = dist[i];
        (*current_correspondences_set_pcl)[i] = pcl::Correspondence((int)i, idx[i], float(dist[i]));
    }
}

void IterativeSE3Registration::update_correspondences_kd_tree_XYZ(const open3d::geometry::KDTreeFlann&) {
    se3icp_ctx* c = context();
    size_t n = source_moving_.points_.size(), m = target_.points_.size();
    if (!c || n == 0 || m == 0) return;
    std::vector<int> idx(n);
    std::vector<double> d2(n);
    int rc = se3icp_nn_xyz(c, xyz_ptr(source_moving_), n, xyz_ptr(target_), m, idx.data(), d2.data());
    if (rc != SE3ICP_OK) return report("se3icp_nn_xyz", rc);
    for (double& v : d2) v = std::sqrt(v);
    store_correspondences(idx, d2);
}

void IterativeSE3Registration::update_correspondences_raw_flann_SE3() {
    update_correspondences_raw_flann_SE3(raw_flann_kd_tree_target_SE3, source_se3_cloud_);
}

void IterativeSE3Registration::update_correspondences_raw_flann_SE3(const open3d::geometry::KDTreeFlann&,
                                                                    const std::vector<Eigen::Matrix4d>& cloud_vector) {
    se3icp_ctx* c = context();
    size_t n = cloud_vector.size(), m = target_se3_cloud_.size();
    if (!c || n == 0 || m == 0) return;
    auto rows_of = [](const std::vector<Eigen::Matrix4d>& v) {
        std::vector<double> rows(12 * v.size());
        for (size_t i = 0; i < v.size(); i++)
            for (int col = 0; col < 4; col++)
                for (int r = 0; r < 3; r++) rows[12 * i + 3 * col + r] = v[i](r, col);
        return rows;
    };
    std::vector<double> rs = rows_of(cloud_vector), rt = rows_of(target_se3_cloud_);
    std::vector<int> idx(n);
    int rc = se3icp_nn_se3(c, rs.data(), n, rt.data(), m, SE3ICP_NN_AUTO, idx.data(), nullptr, nullptr);
    if (rc != SE3ICP_OK) return report("se3icp_nn_se3", rc);
    std::vector<double> dist(n);
    for (size_t i = 0; i < n; i++) {  // reference .cpp:465: 3-D distance of the translation columns
        Eigen::Vector4d d = cloud_vector[i].col(3) - target_se3_cloud_[idx[i]].col(3);
        dist[i] = d.norm();
    }
    store_correspondences(idx, dist);
}

void IterativeSE3Registration::run_entry(int entry, const std::string& variant_name) {
    int variant = entry == SE3ICP_RUN_SE3_ICP_CF ? (int)SE3ICP_GICP : variant_from_name(variant_name);
    if (variant < 0) {
        // reference .cpp:478-480 / :561-563,700-703: message, then run_se3_icp / run_se3_pure leave R = I and,
        // through the un-normalisation at .cpp:735-738, t = c_tgt - c_src.  (run_icp is undefined there.)
        if (entry == SE3ICP_RUN_ICP) {
            std::cerr << "Invalid ICP variant name. Valid names are pt2pt, pt2pl and gicp.\n";
            current_estimated_T_.setIdentity();
            return;
        }
        std::cerr << "Invalid variant name. Choose one of: pt2pt, pt2pl, gicp \n";
        std::cout << "Unknown optimization strategy for SE(3) \n";
        current_estimated_T_.setIdentity();
        Eigen::Vector3d t = target_.GetCenter() - source_.GetCenter();
        current_estimated_T_.block<3, 1>(0, 3) = t;
        num_iterations_ = 1;
        num_pure_se3_iterations_ = 1;
        return;
    }
    se3icp_ctx* c = context();
    if (!c) return;
    int rc = se3icp_set_cloud(c, SE3ICP_SOURCE, xyz_ptr(source_), source_.points_.size(), 0);
    if (rc == SE3ICP_OK) rc = se3icp_set_cloud(c, SE3ICP_TARGET, xyz_ptr(target_), target_.points_.size(), 0);
    if (rc != SE3ICP_OK) return report("se3icp_set_cloud", rc);

    se3icp_params p;
    se3icp_default_params(&p);
    p.variant = variant;
    p.entry = entry;
    p.max_num_iterations = max_num_iterations_;
    p.max_num_se3_iterations = max_num_se3_iterations_;
    p.number_of_nn_for_LRF = number_of_nn_for_LRF_;
    p.trim_keep_largest = trim_keep_largest_ ? 1 : 0;  // default: PCL's comparator (keep largest)
    p.mse = mse_;
    p.mse_switch_error = mse_switch_error_;
    p.estimated_overlap = estimated_overlap_;
    p.alpha_rot = alpha_rot;
    p.beta_transl = beta_transl;
    p.scale_preprocessing = scale_preprocessing;
    p.record_history = entry == SE3ICP_RUN_ICP;
    p.lrf_method = use_shot_lrf_ ? SE3ICP_LRF_SHOT : SE3ICP_LRF_TOLDI;
    p.lrf_radius = lrf_radius_;

    double T[16];
    se3icp_stats st;
    rc = se3icp_run(c, &p, T, &st);
    if (rc != SE3ICP_OK) return report("se3icp_run", rc);

    current_estimated_T_ = from_row_major(T);
    num_iterations_ = st.num_iterations;
    if (entry != SE3ICP_RUN_ICP) {
        num_pure_se3_iterations_ = st.num_pure_se3_iterations;
        time_se3_correspondence_search_ = entry == SE3ICP_RUN_SE3_ICP_CF ? st.time_se3_correspondence_search_ms : 0.0;
    }
    if (entry == SE3ICP_RUN_SE3_ICP_CF) time_before_pure_icp_ = st.time_before_pure_icp_ms;
    if (entry == SE3ICP_RUN_SE3_PURE) std::cout << "pure se3 finished" << std::endl;

    if (entry == SE3ICP_RUN_ICP) {  // reference .cpp:491,538
        estimated_history_.push_back(Eigen::Matrix4d::Identity());
        int cnt = 0;
        std::vector<double> hist((size_t)std::max(max_num_iterations_, 1) * 16);
        if (se3icp_get_history(c, hist.data(), std::max(max_num_iterations_, 1), &cnt) == SE3ICP_OK)
            for (int k = 0; k < cnt && k < std::max(max_num_iterations_, 1); k++)
                estimated_history_.push_back(from_row_major(&hist[16 * (size_t)k]));
    }
    if (mirror_state_) {
        size_t n = source_.points_.size(), m = target_.points_.size();
        // clouds as the reference leaves them: run_icp transforms source_moving_ only (.cpp:541); the SE(3) entries
        // first centre and scale all three clouds in place (.cpp:568-582), then move source_moving_ (.cpp:706)
        Eigen::Matrix4d T_work = current_estimated_T_;
        if (entry != SE3ICP_RUN_ICP) {
            const double s = st.scaling_factor;
            const Eigen::Vector3d cs = source_.GetCenter(), ct = target_.GetCenter();
            const Eigen::Matrix3d R = current_estimated_T_.block<3, 3>(0, 0);
            const Eigen::Vector3d t = current_estimated_T_.block<3, 1>(0, 3);
            T_work.block<3, 1>(0, 3) = (t + R * cs - ct) * s;  // inverse of the un-normalisation at .cpp:735-738
            source_.Translate(-cs);
            source_moving_.Translate(-cs);
            target_.Translate(-ct);
            source_.Scale(s, Eigen::Vector3d(0, 0, 0));
            source_moving_.Scale(s, Eigen::Vector3d(0, 0, 0));
            target_.Scale(s, Eigen::Vector3d(0, 0, 0));
        }
        source_moving_.Transform(T_work);
        std::vector<int> idx(n);
        std::vector<double> dist(n);
        if (se3icp_get_correspondences(c, idx.data(), dist.data(), n) == SE3ICP_OK) store_correspondences(idx, dist);
        if (entry != SE3ICP_RUN_ICP) {
            std::vector<double> fr(16 * std::max(n, m));
            if (se3icp_get_se3_cloud(c, SE3ICP_SOURCE, fr.data(), n) == SE3ICP_OK) {
                source_se3_cloud_.resize(n);
                for (size_t i = 0; i < n; i++) source_se3_cloud_[i] = from_row_major(&fr[16 * i]);
            }
            if (se3icp_get_se3_cloud(c, SE3ICP_TARGET, fr.data(), m) == SE3ICP_OK) {
                target_se3_cloud_.resize(m);
                for (size_t i = 0; i < m; i++) target_se3_cloud_[i] = from_row_major(&fr[16 * i]);
            }
        }
    }
}

void IterativeSE3Registration::run_icp(const std::string& variant_name) { run_entry(SE3ICP_RUN_ICP, variant_name); }
void IterativeSE3Registration::run_se3_icp(const std::string& variant_name) { run_entry(SE3ICP_RUN_SE3_ICP, variant_name); }
void IterativeSE3Registration::run_se3_icp_with_cf() { run_entry(SE3ICP_RUN_SE3_ICP_CF, "gicp"); }
void IterativeSE3Registration::run_se3_pure(const std::string& variant_name) { run_entry(SE3ICP_RUN_SE3_PURE, variant_name); }
