// ceres/gradient_checker.h -- factor/imu_factor.h:20 and projectionTwoFrameOneCamFactor.cpp include it without using it on the hot path;
// with the shim first on the include path it must not pull the real Ceres in beside the shim's ceres:: names.
#pragma once
#include "ceres.h"
