// TEST INFRASTRUCTURE ONLY — never included by the product build (guarded by -DSE3ICP_REFERENCE_BUILD).
//
// CPU restatement of pcl::registration::CorrespondenceRejectorTrimmed (PCL 1.14, un-vendored dependency of the
// reference, README.md:21; call sites src/iterative_SE3_registration.cpp:487-488,508-510,634-635,669-671), restated
// from its published behaviour: keep int(float(N) * overlap_ratio) correspondences selected with std::nth_element on
// the float distance; when that count is >= N the input passes through unchanged, in its original order.
//
// PCL passes pcl::isBetterCorrespondence to nth_element, and that comparator is `pc1.distance > pc2.distance`
// (pcl/correspondence.h), so the LARGEST distances survive.  That source is not in this image, so the direction is
// restated, not verified: `trim_keep_largest()` selects it, default 1 = PCL's comparator, the same
// default and the same switch as oracle/se3icp_oracle.cpp and include/se3icp.h.  Equal distances at the cut resolve to the smaller query index (nth_element leaves that open).
#pragma once
#include <algorithm>
#include <vector>

#include "../../compat/pcl/correspondence.h"

namespace pcl {
namespace registration {

inline int& trim_keep_largest() {
    static int flag = 1;
    return flag;
}

class CorrespondenceRejectorTrimmed {
public:
    void setOverlapRatio(float ratio) { overlap_ratio_ = std::min(1.0f, std::max(0.0f, ratio)); }
    float getOverlapRatio() const { return overlap_ratio_; }
    void setMinCorrespondences(unsigned int n) { nr_min_correspondences_ = n; }
    void setInputCorrespondences(const CorrespondencesConstPtr& c) { input_ = c; }
    void getCorrespondences(Correspondences& out) const {
        if (!input_) {
            out.clear();
            return;
        }
        getRemainingCorrespondences(*input_, out);
    }
    void getRemainingCorrespondences(const Correspondences& original, Correspondences& remaining) const {
        unsigned int keep = (unsigned int)(int)((float)original.size() * overlap_ratio_);
        keep = std::max(keep, nr_min_correspondences_);
        remaining = original;
        if (keep >= original.size()) return;
        const bool largest = trim_keep_largest() != 0;
        std::nth_element(remaining.begin(), remaining.begin() + keep, remaining.end(),
                         [largest](const Correspondence& a, const Correspondence& b) {
                             if (a.distance != b.distance) return largest ? a.distance > b.distance : a.distance < b.distance;
                             return a.index_query < b.index_query;
                         });
        remaining.resize(keep);
    }

private:
    float overlap_ratio_ = 0.5f;
    unsigned int nr_min_correspondences_ = 0;
    CorrespondencesConstPtr input_;
};

}  // namespace registration
}  // namespace pcl
