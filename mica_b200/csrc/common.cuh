// Shared helpers for the mica_b200 kernels (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/mica_b200.h"

namespace mica {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

extern thread_local char g_last_error[512];
extern std::atomic<int64_t> g_launches;

int set_error(int code, const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);

// How long a kernel waits for a flag another rank raises over NVLink before it gives up with a status word
// (clock64 ticks).  Ranks of one job can be SECONDS apart -- one still allocating 10 GB of pinned host memory,
// another draining its volumes over a shared PCIe switch -- so the bound only has to beat "forever":
// MICA_PEER_TIMEOUT_MS, default 60 s.
long long peer_timeout_cycles();

inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define MICA_CUDA(call)                                      \
  do {                                                       \
    int _rc = ::mica::check_cuda((call), #call);             \
    if (_rc != MICA_OK) return _rc;                          \
  } while (0)

#define MICA_LAUNCH_CHECK(name)                              \
  do {                                                       \
    ::mica::count_launch();                                  \
    int _rc = ::mica::check_cuda(cudaGetLastError(), name);  \
    if (_rc != MICA_OK) return _rc;                          \
  } while (0)

#define MICA_REQUIRE(cond, ...)                                            \
  do {                                                                     \
    if (!(cond)) return ::mica::set_error(MICA_ERR_INVALID, __VA_ARGS__);  \
  } while (0)

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda at link time)
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TensorMapEncodeFn tensor_map_encode_fn();

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// streaming (read-once / write-once) accesses: keep L1 for the data with reuse
__device__ __forceinline__ float ld_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
// same, for 128-byte runs that start in the middle of a 128-byte line (cube cores at stride 32):
// by default a miss makes L2 fetch the whole 128-byte line from DRAM (measured on B200: 2.0x the
// requested bytes, tools/micro/overfetch.cu); .L2::64B caps the fill at the 64-byte half that is used
__device__ __forceinline__ float ld_stream_half_line(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.L2::64B.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ld_stream4(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream(float* p, float v) {
  asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void st_stream4(float4* p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// system-scope release / acquire on a flag word in (possibly peer) global memory: the PTX-model form of
// "data, fence, flag" between GPUs
__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

}  // namespace mica
