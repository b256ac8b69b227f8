// stand-in for <pcl/registration/correspondence_rejection_trimmed.h> (reference hpp:7): the trimmed rejection runs on the GPU
#pragma once
#include "../correspondence.h"
