// empty shim: lib/utils.h includes <pcl/point_cloud.h> but uses nothing from it (oracle/_ref build only)
#pragma once
#include <cmath>
