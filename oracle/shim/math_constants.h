// oracle/shim/math_constants.h — TEST INFRASTRUCTURE: the one constant the reference takes from
// CUDA's math_constants.h (camera.cuh:62, sphere.cuh:45-48).
#pragma once
#define CUDART_PI_F 3.141592654f
