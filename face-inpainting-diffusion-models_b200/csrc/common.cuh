// Shared helpers for the fidm_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/fidm_b200.h"

namespace fidm {

void set_error(const char* fmt, ...);

#define FIDM_REQUIRE(cond, code, ...)                   \
  do {                                                  \
    if (!(cond)) {                                      \
      ::fidm::set_error(__VA_ARGS__);                   \
      return (code);                                    \
    }                                                   \
  } while (0)

// Call after a kernel launch.  Returns a cudaError_t (>0) through the C ABI.
#define FIDM_CHECK_LAUNCH(what)                                                    \
  do {                                                                             \
    cudaError_t e__ = cudaGetLastError();                                          \
    if (e__ != cudaSuccess) {                                                      \
      ::fidm::set_error("%s: %s", (what), cudaGetErrorString(e__));                \
      return (int)e__;                                                             \
    }                                                                              \
  } while (0)

#define FIDM_CUDA(call)                                                            \
  do {                                                                             \
    cudaError_t e__ = (call);                                                      \
    if (e__ != cudaSuccess) {                                                      \
      ::fidm::set_error("%s: %s", #call, cudaGetErrorString(e__));                 \
      return (int)e__;                                                             \
    }                                                                              \
  } while (0)

inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }

// A vector of V elements of T, loaded/stored with one instruction when it is 4/8/16 bytes.
template <typename T, int V> struct alignas(sizeof(T) * V) Vec { T v[V]; };

template <typename T, int V>
__device__ __forceinline__ void load_vec(const T* p, float (&out)[V]) {
  Vec<T, V> t = *reinterpret_cast<const Vec<T, V>*>(p);
#pragma unroll
  for (int i = 0; i < V; ++i) out[i] = to_f32<T>(t.v[i]);
}
template <typename T, int V>
__device__ __forceinline__ void store_vec(T* p, const float (&in)[V]) {
  Vec<T, V> t;
#pragma unroll
  for (int i = 0; i < V; ++i) t.v[i] = from_f32<T>(in[i]);
  *reinterpret_cast<Vec<T, V>*>(p) = t;
}

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

}  // namespace fidm
