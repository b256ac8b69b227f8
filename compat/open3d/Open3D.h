// Minimal stand-in for the slice of Open3D 0.19 the SE(3)-ICP class API and the reference's
// run_registration_method driver touch (reference include/iterative_SE3_registration.hpp:14,20,33-38,
// 54-56,66-68,76-78; examples/run_registration_method.cpp:27-31).  NOT Open3D.  Used only when the
// real library is not installed; the numeric work of the reference's Open3D calls
// is done by libse3icp_cuda.so, not here.
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "../Eigen/Core"

namespace open3d {

namespace geometry {

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
};

// member type only: the spatial index lives on the GPU
class KDTreeFlann {
public:
    KDTreeFlann() = default;
    bool SetGeometry(const PointCloud&) { return true; }
};

}  // namespace geometry

namespace pipelines {
namespace registration {
typedef std::vector<Eigen::Vector2i> CorrespondenceSet;
// member types only: the estimators run inside libse3icp_cuda.so
class TransformationEstimationPointToPoint {};
class TransformationEstimationPointToPlane {};
class TransformationEstimationForGeneralizedICP {};
}  // namespace registration
}  // namespace pipelines

namespace utility {
namespace random {
inline void Seed(int) {}
}  // namespace random
}  // namespace utility

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
