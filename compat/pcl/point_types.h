// stand-in for <pcl/point_types.h> (reference hpp:11): nothing from it is used by the class API
#pragma once
