// Shared helpers for libvidar_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/vidar_b200.h"

namespace vidar {

// thread-local last-error text (vidar_last_error)
char* err_buf();
int set_error(int code, const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

inline int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess)
    return set_error(VIDAR_E_CUDA, "%s: launch failed: %s", what, cudaGetErrorString(e));
  return VIDAR_OK;
}

#define VIDAR_REQUIRE(cond, ...)                                   \
  do {                                                             \
    if (!(cond)) return ::vidar::set_error(VIDAR_E_INVALID, __VA_ARGS__); \
  } while (0)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

__device__ __forceinline__ float4 ldg4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}

// red.global.add.v4.f32 (sm_90+): one 16-byte vector reduction instead of four scalars.
__device__ __forceinline__ void red_add_v4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

__device__ __forceinline__ void red_add_f32(float* p, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

}  // namespace vidar
