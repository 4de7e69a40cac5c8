// tcgen05 / TMEM / TMA kernels (sm_100a) -- placeholder interface, filled in below.
#pragma once
#include <cuda_runtime.h>
#include "kernels_f32.cuh"

namespace b2c {
struct TcWeight { void* hi = nullptr; void* lo = nullptr; };
struct TcConvPlan { int dummy = 0; };
inline void tc_weight_free(TcWeight&) {}
inline int tc_weight_pack(const float*, int, int, int, int, TcWeight*, size_t*) { return 0; }
inline int tc_conv_plan(const ConvArgs&, const TcWeight&, int, int, TcConvPlan*) { return 1; }
inline int tc_conv_launch(const TcConvPlan&, const ConvArgs&, const TcWeight&, cudaStream_t) { return -1; }
inline int tc_nearest_launch(const RvqArgs&, int, int, cudaStream_t) { return 1; }
}  // namespace b2c
