// Forwarding shim (oracle/_ref build, test infrastructure): see gsref_ros.h / gsref_cv.h
#pragma once
#include "../gsref_ros.h"
