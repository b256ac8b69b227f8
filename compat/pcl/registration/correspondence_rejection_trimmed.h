// stand-in for <pcl/registration/correspondence_rejection_trimmed.h> (reference hpp:7): in the product build the trimmed
// rejection runs on the GPU and only the header name is needed.  The oracle's build of the reference source
// (-DSE3ICP_REFERENCE_BUILD) gets a CPU restatement of the PCL 1.14 class from oracle/refdeps/.
#pragma once
#include "../correspondence.h"
#ifdef SE3ICP_REFERENCE_BUILD
#include "../../../oracle/refdeps/pcl_trimmed_rejector.h"
#endif
