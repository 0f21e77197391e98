// Shared device/host helpers for the mtl-ssl B200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

typedef __nv_bfloat16 bf16;

#define MTL_OK 0
#define MTL_ERR_ARG (-1)
#define MTL_ERR_CUDA (-2)
#define MTL_ERR_UNSUPPORTED (-3)

extern "C" void mtl_set_error(const char* fmt, ...);
// every kernel launch of the library passes through MTL_CUDA_LAUNCH_CHECK, which counts it
extern "C" void mtl_count_launch(void);

#define MTL_CHECK_ARG(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      mtl_set_error(__VA_ARGS__);                \
      return MTL_ERR_ARG;                        \
    }                                            \
  } while (0)

#define MTL_CUDA_LAUNCH_CHECK(name)                                          \
  do {                                                                       \
    mtl_count_launch();                                                      \
    cudaError_t e__ = cudaGetLastError();                                    \
    if (e__ != cudaSuccess) {                                                \
      mtl_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
      return MTL_ERR_CUDA;                                                   \
    }                                                                        \
  } while (0)

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// Number of SMs on the current device (cached).
int mtl_num_sms();

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Activation element type helpers: kernels are templated on T in {float, bf16}.
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f32<bf16>(float v) { return __float2bfloat16_rn(v); }
