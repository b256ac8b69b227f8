// Minimal stand-in for pcl::Correspondence(s) (PCL 1.14 common/include/pcl/correspondence.h), the only PCL
// types the class API exposes (reference hpp:40-43,75).  NOT PCL.
#pragma once
#include <memory>
#include <vector>

namespace pcl {
using index_t = int;
struct Correspondence {
    index_t index_query = 0;
    index_t index_match = -1;
    float distance = 3.4028235e38f;
    Correspondence() = default;
    Correspondence(index_t q, index_t m, float d) : index_query(q), index_match(m), distance(d) {}
};
using Correspondences = std::vector<Correspondence>;
using CorrespondencesPtr = std::shared_ptr<Correspondences>;
using CorrespondencesConstPtr = std::shared_ptr<const Correspondences>;
}  // namespace pcl
