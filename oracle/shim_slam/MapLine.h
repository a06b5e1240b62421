// ORACLE shim (test infrastructure): see slam_mock.h
#pragma once
#include "slam_mock.h"
