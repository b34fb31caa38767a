// Shared device helpers for the LightGCN sm_100a kernels.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/lgcn_b200.h"

namespace lgcn {

// ---- error plumbing (no exceptions across the C ABI) ----------------------
void set_last_error(const char* fmt, ...);

#define LGCN_CHECK_ARG(cond, ...)                \
  do {                                           \
    if (!(cond)) {                               \
      ::lgcn::set_last_error(__VA_ARGS__);       \
      return LGCN_ERR_INVALID_ARG;               \
    }                                            \
  } while (0)

#define LGCN_CUDA_OK(expr)                                                            \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      ::lgcn::set_last_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),  \
                             __FILE__, __LINE__);                                     \
      return (int)_e;                                                                 \
    }                                                                                 \
  } while (0)

#define LGCN_LAUNCH_OK()                                                              \
  do {                                                                                \
    cudaError_t _e = cudaGetLastError();                                              \
    if (_e != cudaSuccess) {                                                          \
      ::lgcn::set_last_error("kernel launch failed: %s (%s:%d)",                      \
                             cudaGetErrorString(_e), __FILE__, __LINE__);             \
      return (int)_e;                                                                 \
    }                                                                                 \
  } while (0)

constexpr int kSmCount = 148;  // B200: 2 dies x 74 SMs

// ---- 16-byte vector loads / stores ----------------------------------------
// Gathered embedding rows: read-only path, keep in L1 (popular rows repeat).
__device__ __forceinline__ uint4 ldg_row16(const void* p) {
  return __ldg(reinterpret_cast<const uint4*>(p));
}
// Streaming int loads (column indices): do not pollute L1.
__device__ __forceinline__ int ldg_stream_i32(const int32_t* p) {
  int v;
  asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ldg_f4_stream(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ld_f4(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
__device__ __forceinline__ void st_f4(float* p, float4 v) {
  *reinterpret_cast<float4*>(p) = v;
}
__device__ __forceinline__ void st_u4(void* p, uint4 v) {
  *reinterpret_cast<uint4*>(p) = v;
}

// fp32 vector atomic add (sm_90+: one 16-byte red instead of four).
__device__ __forceinline__ void red_add_f4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b),
               "f"(c), "f"(d)
               : "memory");
}

// ---- bf16 pack / unpack ------------------------------------------------------
__device__ __forceinline__ void unpack_bf16x8(const uint4& v, float (&f)[8]) {
  const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint4 pack_bf16x8(const float (&f)[8]) {
  uint4 v;
  v.x = pack_bf16x2(f[0], f[1]);
  v.y = pack_bf16x2(f[2], f[3]);
  v.z = pack_bf16x2(f[4], f[5]);
  v.w = pack_bf16x2(f[6], f[7]);
  return v;
}

// ---- reductions ------------------------------------------------------------
template <int WIDTH>
__device__ __forceinline__ float group_sum(float v, unsigned mask) {
#pragma unroll
  for (int o = WIDTH / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o, WIDTH);
  return v;
}

__device__ __forceinline__ unsigned group_mask(int lanes_per_group) {
  // mask of the lanes of the caller's group inside its warp
  const int lane = threadIdx.x & 31;
  if (lanes_per_group >= 32) return 0xffffffffu;
  const unsigned base = (1u << lanes_per_group) - 1u;
  return base << ((lane / lanes_per_group) * lanes_per_group);
}

}  // namespace lgcn
