// ORACLE shim (test infrastructure): the mocks of ../shim_slam/slam_mock.h WITHOUT the mock Frame - this build compiles the reference's real Frame.h / Frame.cc
#pragma once
#define PL_SHIM_REAL_FRAME 1
#include "../shim_slam/slam_mock.h"
