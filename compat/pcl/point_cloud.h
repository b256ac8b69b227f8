// stand-in for <pcl/point_cloud.h> (reference hpp:10): nothing from it is used by the class API
#pragma once
