// Minimal stand-in for the slice of Open3D 0.19 the SE(3)-ICP class API and the reference's
// run_registration_method driver touch (reference include/iterative_SE3_registration.hpp:14,20,33-38,
// 54-56,66-68,76-78; examples/run_registration_method.cpp:27-31).  NOT Open3D.  Used only when the
// real library is not installed; the numeric work of the reference's Open3D calls
// is done by libse3icp_cuda.so, not here.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <numeric>
#include <random>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "../Eigen/Core"

namespace open3d {

namespace utility {
namespace random {
inline std::mt19937& Engine() {
    static std::mt19937 e(0);
    return e;
}
inline void Seed(int seed) { Engine().seed((unsigned)seed); }
}  // namespace random
typedef std::allocator<Eigen::Vector6d> Vector6d_allocator;
template <typename... Args>
inline void LogDebug(const char*, Args&&...) {}
inline Eigen::Matrix3d SkewMatrix(const Eigen::Vector3d& v) {
    Eigen::Matrix3d m;
    m(0, 1) = -v[2], m(0, 2) = v[1], m(1, 0) = v[2], m(1, 2) = -v[0], m(2, 0) = -v[1], m(2, 1) = v[0];
    return m;
}
}  // namespace utility

namespace geometry {

class KDTreeSearchParam {
public:
    virtual ~KDTreeSearchParam() {}
};
class KDTreeSearchParamKNN : public KDTreeSearchParam {
public:
    KDTreeSearchParamKNN(int knn = 30) : knn_(knn) {}
    int knn_;
};

class PointCloud {
public:
    std::vector<Eigen::Vector3d> points_;
    std::vector<Eigen::Vector3d> normals_;
    std::vector<Eigen::Matrix3d> covariances_;

    bool HasPoints() const { return !points_.empty(); }
    bool HasNormals() const { return !points_.empty() && normals_.size() == points_.size(); }
    bool HasCovariances() const { return !points_.empty() && covariances_.size() == points_.size(); }

    Eigen::Vector3d GetCenter() const {
        Eigen::Vector3d c;
        if (points_.empty()) return c;
        for (const auto& p : points_) c += p;
        return c / (double)points_.size();
    }
    PointCloud& Translate(const Eigen::Vector3d& t, bool relative = true) {
        Eigen::Vector3d shift = t;
        if (!relative) shift = t - GetCenter();
        for (auto& p : points_) p += shift;
        return *this;
    }
    PointCloud& Scale(double s, const Eigen::Vector3d& center) {
        for (auto& p : points_) p = (p - center) * s + center;
        return *this;
    }
    PointCloud& Transform(const Eigen::Matrix4d& T) {
        Eigen::Matrix3d R = T.block<3, 3>(0, 0);
        for (auto& p : points_) {
            Eigen::Vector4d h = T * Eigen::Vector4d(p[0], p[1], p[2], 1.0);
            p = Eigen::Vector3d(h[0] / h[3], h[1] / h[3], h[2] / h[3]);
        }
        for (auto& n : normals_) n = R * n;
        for (auto& C : covariances_) C = R * C * R.transpose();
        return *this;
    }
    // random subset of round(ratio * n) points, seeded through utility::random::Seed (the index stream differs
    // from Open3D's, the statistics do not)
    std::shared_ptr<PointCloud> RandomDownSample(double sampling_ratio) const {
        auto out = std::make_shared<PointCloud>();
        std::vector<size_t> idx(points_.size());
        std::iota(idx.begin(), idx.end(), 0);
        std::shuffle(idx.begin(), idx.end(), utility::random::Engine());
        size_t n = (size_t)((double)points_.size() * sampling_ratio);
        idx.resize(std::min(n, idx.size()));
        for (size_t i : idx) out->points_.push_back(points_[i]);
        return out;
    }
    inline void EstimateNormals(const KDTreeSearchParam& search_param = KDTreeSearchParamKNN(),
                                bool fast_normal_computation = true);
    inline std::vector<double> ComputePointCloudDistance(const PointCloud& target) const;
    PointCloud& PaintUniformColor(const Eigen::Vector3d&) { return *this; }
};

// The registration path never searches this tree (its spatial indices live on the GPU); the stand-in is a plain
// CPU kd-tree of any dimension (SetGeometry: 3-D, SetMatrixData: one point per COLUMN) so that the reference's
// evaluation helpers (src/cc.cpp:116-143,220-237) work.  Exact search; equal distances resolve to the smallest index.
class KDTreeFlann {
public:
    KDTreeFlann() = default;
    explicit KDTreeFlann(const PointCloud& pc) { SetGeometry(pc); }
    bool SetGeometry(const PointCloud& pc) {
        dim_ = 3;
        data_.resize(pc.points_.size() * 3);
        for (size_t i = 0; i < pc.points_.size(); i++)
            for (int d = 0; d < 3; d++) data_[i * 3 + d] = pc.points_[i][d];
        return Rebuild();
    }
    bool SetMatrixData(const Eigen::MatrixXd& data) {
        dim_ = (int)data.rows();
        data_.assign(data.data(), data.data() + data.size());  // column-major: column i = point i
        return Rebuild();
    }
    template <typename Q>
    int SearchKNN(const Q& q, int knn, std::vector<int>& indices, std::vector<double>& distance2) const {
        std::vector<std::pair<double, int>> heap;
        double qv[kMaxDim];
        for (int d = 0; d < dim_; d++) qv[d] = q(d);
        if (!nodes_.empty() && knn > 0) Search(0, qv, (size_t)knn, 1e300, heap);
        std::sort_heap(heap.begin(), heap.end());
        indices.resize(heap.size());
        distance2.resize(heap.size());
        for (size_t i = 0; i < heap.size(); i++) {
            indices[i] = heap[i].second;
            distance2[i] = heap[i].first;
        }
        return (int)heap.size();
    }
    template <typename Q>
    int SearchRadius(const Q& q, double radius, std::vector<int>& indices, std::vector<double>& distance2) const {
        std::vector<std::pair<double, int>> heap;
        double qv[kMaxDim];
        for (int d = 0; d < dim_; d++) qv[d] = q(d);
        if (!nodes_.empty()) Search(0, qv, n_points(), radius * radius, heap);
        std::sort_heap(heap.begin(), heap.end());
        indices.clear();
        distance2.clear();
        for (auto& h : heap) {
            if (!(h.first < radius * radius)) continue;  // nanoflann's RadiusResultSet::addPoint keeps dist < radius^2 (strict)
            indices.push_back(h.second);
            distance2.push_back(h.first);
        }
        return (int)indices.size();
    }

private:
    static constexpr int kMaxDim = 16;
    struct Node {
        int left = -1, right = -1, dim = -1, begin = 0, end = 0;
        double split = 0;
    };
    int dim_ = 3;
    std::vector<double> data_;
    std::vector<int> perm_;
    std::vector<Node> nodes_;
    size_t n_points() const { return data_.size() / (size_t)dim_; }
    double at(int i, int d) const { return data_[(size_t)i * dim_ + d]; }
    bool Rebuild() {
        if (dim_ < 1 || dim_ > kMaxDim) return false;
        perm_.resize(n_points());
        std::iota(perm_.begin(), perm_.end(), 0);
        nodes_.clear();
        if (!perm_.empty()) Build(0, (int)perm_.size());
        return true;
    }
    int Build(int b, int e) {
        int id = (int)nodes_.size();
        nodes_.emplace_back();
        nodes_[id].begin = b;
        nodes_[id].end = e;
        if (e - b <= 16) return id;
        int dim = 0;
        double best = -1.0;
        for (int d = 0; d < dim_; d++) {
            double lo = 1e300, hi = -1e300;
            for (int i = b; i < e; i++) lo = std::min(lo, at(perm_[i], d)), hi = std::max(hi, at(perm_[i], d));
            if (hi - lo > best) best = hi - lo, dim = d;
        }
        int mid = (b + e) / 2;
        std::nth_element(perm_.begin() + b, perm_.begin() + mid, perm_.begin() + e,
                         [&](int x, int y) { return at(x, dim) < at(y, dim); });
        nodes_[id].dim = dim;
        nodes_[id].split = at(perm_[mid], dim);
        int l = Build(b, mid);
        int r = Build(mid, e);
        nodes_[id].left = l;
        nodes_[id].right = r;
        return id;
    }
    void Search(int id, const double* q, size_t k, double max_d2, std::vector<std::pair<double, int>>& heap) const {
        const Node& nd = nodes_[id];
        if (nd.dim < 0) {
            for (int i = nd.begin; i < nd.end; i++) {
                double d2 = 0.0;
                for (int d = 0; d < dim_; d++) {
                    double t = at(perm_[i], d) - q[d];
                    d2 += t * t;
                }
                if (d2 > max_d2) continue;
                std::pair<double, int> cand(d2, perm_[i]);
                if (heap.size() < k) {
                    heap.push_back(cand);
                    std::push_heap(heap.begin(), heap.end());
                } else if (cand < heap.front()) {
                    std::pop_heap(heap.begin(), heap.end());
                    heap.back() = cand;
                    std::push_heap(heap.begin(), heap.end());
                }
            }
            return;
        }
        double diff = q[nd.dim] - nd.split;
        int near = diff < 0 ? nd.left : nd.right, far = diff < 0 ? nd.right : nd.left;
        Search(near, q, k, max_d2, heap);
        double bound = heap.size() < k ? max_d2 : std::min(max_d2, heap.front().first);
        if (diff * diff <= bound) Search(far, q, k, max_d2, heap);
    }
};

inline std::vector<double> PointCloud::ComputePointCloudDistance(const PointCloud& target) const {
    std::vector<double> out(points_.size());
    KDTreeFlann tree(target);
    std::vector<int> idx;
    std::vector<double> d2;
    for (size_t i = 0; i < points_.size(); i++) out[i] = tree.SearchKNN(points_[i], 1, idx, d2) ? std::sqrt(d2[0]) : 0.0;
    return out;
}

}  // namespace geometry

namespace pipelines {
namespace registration {
typedef std::vector<Eigen::Vector2i> CorrespondenceSet;
// Member types of the registration class.  In the product build the estimators run inside libse3icp_cuda.so and
// ComputeTransformation is declared but never defined or called; the CPU definitions exist only in the oracle's
// build of the reference source (oracle/refdeps/third_party_numerics.h, -DSE3ICP_REFERENCE_BUILD).
class TransformationEstimationPointToPoint {
public:
    Eigen::Matrix4d ComputeTransformation(const geometry::PointCloud& source, const geometry::PointCloud& target,
                                          const CorrespondenceSet& corres) const;
};
class TransformationEstimationPointToPlane {
public:
    Eigen::Matrix4d ComputeTransformation(const geometry::PointCloud& source, const geometry::PointCloud& target,
                                          const CorrespondenceSet& corres) const;
};
class TransformationEstimationForGeneralizedICP {
public:
    Eigen::Matrix4d ComputeTransformation(const geometry::PointCloud& source, const geometry::PointCloud& target,
                                          const CorrespondenceSet& corres) const;
};

// FPFH + Fast Global Registration are a different algorithm (baseline branches of the reference's drivers,
// OUT OF SCOPE in SURVEY §2.1).  The stand-ins let those translation units compile; calling them reports it.
class Feature {};
class FastGlobalRegistrationOption {};
class RegistrationResult {
public:
    Eigen::Matrix4d transformation_ = Eigen::Matrix4d::Identity();
    double fitness_ = 0.0, inlier_rmse_ = 0.0;
};
inline std::shared_ptr<Feature> ComputeFPFHFeature(const geometry::PointCloud&, const geometry::KDTreeSearchParam& =
                                                                                    geometry::KDTreeSearchParamKNN()) {
    std::cerr << "[compat] FPFH features are not provided by the stand-in (FGR baseline is out of scope)\n";
    return std::make_shared<Feature>();
}
inline RegistrationResult FastGlobalRegistrationBasedOnFeatureMatching(const geometry::PointCloud&, const geometry::PointCloud&,
                                                                       const Feature&, const Feature&,
                                                                       const FastGlobalRegistrationOption& =
                                                                           FastGlobalRegistrationOption()) {
    std::cerr << "[compat] Fast Global Registration is not provided by the stand-in; returning identity\n";
    return RegistrationResult();
}
}  // namespace registration
}  // namespace pipelines

namespace visualization {
inline bool DrawGeometries(const std::vector<std::shared_ptr<const geometry::PointCloud>>&, const std::string& = "Open3D",
                           int = 640, int = 480, int = 50, int = 50) {
    std::cerr << "[compat] open3d::visualization is not provided by the stand-in\n";
    return false;
}
}  // namespace visualization

namespace io {

// PLY reader: ascii or binary_little_endian, vertex element with x/y/z as float or double; every
// other property and element is skipped.  Returns false (and leaves the cloud untouched) on failure,
// as open3d::io::ReadPointCloud does.
inline bool ReadPointCloud(const std::string& filename, geometry::PointCloud& cloud) {
    std::ifstream f(filename, std::ios::binary);
    if (!f) {
        std::cerr << "[Open3D WARNING] Read PLY failed: unable to open file: " << filename << std::endl;
        return false;
    }
    std::string line;
    std::getline(f, line);
    if (line.substr(0, 3) != "ply") return false;
    struct Prop {
        std::string name, type;
        bool list = false;
        std::string count_type;
    };
    struct Elem {
        std::string name;
        size_t count = 0;
        std::vector<Prop> props;
    };
    std::vector<Elem> elems;
    std::string fmt;
    while (std::getline(f, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        std::istringstream ss(line);
        std::string tok;
        ss >> tok;
        if (tok == "format") {
            ss >> fmt;
        } else if (tok == "element") {
            Elem e;
            ss >> e.name >> e.count;
            elems.push_back(e);
        } else if (tok == "property" && !elems.empty()) {
            Prop p;
            ss >> p.type;
            if (p.type == "list") {
                p.list = true;
                ss >> p.count_type >> p.type >> p.name;
            } else {
                ss >> p.name;
            }
            elems.back().props.push_back(p);
        } else if (tok == "end_header") {
            break;
        }
    }
    auto size_of = [](const std::string& t) -> int {
        if (t == "char" || t == "uchar" || t == "int8" || t == "uint8") return 1;
        if (t == "short" || t == "ushort" || t == "int16" || t == "uint16") return 2;
        if (t == "int" || t == "uint" || t == "float" || t == "int32" || t == "uint32" || t == "float32") return 4;
        if (t == "double" || t == "float64") return 8;
        return 0;
    };
    const bool ascii = fmt == "ascii";
    if (!ascii && fmt != "binary_little_endian") return false;
    std::vector<Eigen::Vector3d> pts;
    for (const Elem& e : elems) {
        const bool is_vertex = e.name == "vertex";
        if (is_vertex) pts.reserve(e.count);
        for (size_t i = 0; i < e.count; i++) {
            Eigen::Vector3d p;
            for (const Prop& pr : e.props) {
                if (pr.list) {
                    long cnt = 0;
                    if (ascii) {
                        f >> cnt;
                        double skip;
                        for (long k = 0; k < cnt; k++) f >> skip;
                    } else {
                        unsigned char buf[8] = {0};
                        f.read((char*)buf, size_of(pr.count_type));
                        cnt = buf[0] | (buf[1] << 8) | (buf[2] << 16) | ((long)buf[3] << 24);
                        f.ignore(cnt * size_of(pr.type));
                    }
                    continue;
                }
                double v = 0.0;
                if (ascii) {
                    f >> v;
                } else {
                    char buf[8];
                    int sz = size_of(pr.type);
                    f.read(buf, sz);
                    if (pr.type == "float" || pr.type == "float32") {
                        float t;
                        std::memcpy(&t, buf, 4);
                        v = t;
                    } else if (pr.type == "double" || pr.type == "float64") {
                        std::memcpy(&v, buf, 8);
                    }
                }
                if (is_vertex) {
                    if (pr.name == "x") p[0] = v;
                    if (pr.name == "y") p[1] = v;
                    if (pr.name == "z") p[2] = v;
                }
            }
            if (is_vertex) pts.push_back(p);
        }
        if (is_vertex) break;  // nothing after the vertices is needed
    }
    if (!f && !f.eof()) return false;
    cloud.points_ = pts;
    cloud.normals_.clear();
    cloud.covariances_.clear();
    return true;
}

inline std::shared_ptr<geometry::PointCloud> CreatePointCloudFromFile(const std::string& filename) {
    auto cloud = std::make_shared<geometry::PointCloud>();
    ReadPointCloud(filename, *cloud);
    return cloud;
}

inline bool WritePointCloud(const std::string& filename, const geometry::PointCloud& cloud) {
    std::ofstream f(filename, std::ios::binary);
    if (!f) return false;
    f << "ply\nformat binary_little_endian 1.0\ncomment Created by se3-icp_b200 compat\nelement vertex " << cloud.points_.size()
      << "\nproperty double x\nproperty double y\nproperty double z\nend_header\n";
    for (const auto& p : cloud.points_) f.write((const char*)p.data(), 3 * sizeof(double));
    return (bool)f;
}

}  // namespace io
}  // namespace open3d

#ifdef SE3ICP_REFERENCE_BUILD
// oracle-only: CPU restatements of the third-party numerics the reference source calls (test infrastructure)
#include "../../oracle/refdeps/third_party_numerics.h"
#else
// Only the FGR baselines and diagnostics of the reference's drivers call this (out of scope);
// the registration path computes its normals on the GPU.
inline void open3d::geometry::PointCloud::EstimateNormals(const KDTreeSearchParam&, bool) {
    std::cerr << "[compat] open3d::geometry::PointCloud::EstimateNormals is not provided by the stand-in\n";
}
#endif
