// Microbenchmark: DRAM bytes fetched when reading 128-byte runs that start 64 bytes into a
// 256-byte row (the stride-32 core of a 64^3 logit cube), with different load flavours.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o overfetch overfetch.cu
// Run under: ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum ./overfetch
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE> __device__ __forceinline__ float ld(const float* p) {
  float v;
  if (MODE == 0) asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  if (MODE == 1) asm volatile("ld.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
  if (MODE == 2) asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(v) : "l"(p));
  if (MODE == 3) asm volatile("ld.global.cs.f32 %0, [%1];" : "=f"(v) : "l"(p));
  if (MODE == 4) asm volatile("ld.global.L2::64B.f32 %0, [%1];" : "=f"(v) : "l"(p));
  if (MODE == 5) asm volatile("ld.global.L2::128B.f32 %0, [%1];" : "=f"(v) : "l"(p));
  if (MODE == 6) asm volatile("ld.global.lu.f32 %0, [%1];" : "=f"(v) : "l"(p));
  if (MODE == 7) {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("ld.global.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
  }
  return v;
}

// grid = (32 planes, B*C), block 256; reads [off, off+32) floats of rows 16..47 of planes 16..47
template <int MODE>
__global__ void core_read(const float* __restrict__ in, float* __restrict__ out, int off) {
  const float* cube = in + (size_t)blockIdx.y * 262144 + (size_t)(blockIdx.x + 16) * 4096;
  float acc = 0.f;
  for (int e = threadIdx.x; e < 1024; e += 256) {
    int r = e >> 5, c = e & 31;
    acc += ld<MODE>(cube + (r + 16) * 64 + off + c);
  }
  if (acc == 123.456f) out[0] = acc;
}

// float4 flavour: 8 lanes cover one 128-byte run
__global__ void core_read_v4(const float* __restrict__ in, float* __restrict__ out, int off) {
  const float* cube = in + (size_t)blockIdx.y * 262144 + (size_t)(blockIdx.x + 16) * 4096;
  float acc = 0.f;
  int r = threadIdx.x >> 3, c = (threadIdx.x & 7) * 4;
  float4 v = *reinterpret_cast<const float4*>(cube + (r + 16) * 64 + off + c);
  acc = v.x + v.y + v.z + v.w;
  if (acc == 123.456f) out[0] = acc;
}

// whole rows (256 B) of the core planes/rows: the "just read both lines" baseline
__global__ void rows_read(const float* __restrict__ in, float* __restrict__ out) {
  const float* cube = in + (size_t)blockIdx.y * 262144 + (size_t)(blockIdx.x + 16) * 4096;
  float acc = 0.f;
  for (int e = threadIdx.x; e < 2048; e += 256) {
    int r = e >> 6, c = e & 63;
    acc += cube[(r + 16) * 64 + c];
  }
  if (acc == 123.456f) out[0] = acc;
}

int main() {
  const int BC = 32 * 29;   // 928 cubes of 1 MiB = 0.97 GB, far above L2
  float *in, *out;
  cudaMalloc(&in, (size_t)BC * 262144 * 4);
  cudaMalloc(&out, 4);
  cudaMemset(in, 0, (size_t)BC * 262144 * 4);
  dim3 grid(32, BC);
  printf("algorithmic bytes per launch: %.1f MB\n", BC * 32.0 * 32 * 128 / 1e6);
  for (int off = 16; off >= 0; off -= 16) {
    core_read<0><<<grid, 256>>>(in, out, off);
    core_read<1><<<grid, 256>>>(in, out, off);
    core_read<2><<<grid, 256>>>(in, out, off);
    core_read<3><<<grid, 256>>>(in, out, off);
    core_read<4><<<grid, 256>>>(in, out, off);
    core_read<5><<<grid, 256>>>(in, out, off);
    core_read<6><<<grid, 256>>>(in, out, off);
    core_read<7><<<grid, 256>>>(in, out, off);
    core_read_v4<<<grid, 256>>>(in, out, off);
  }
  rows_read<<<grid, 256>>>(in, out);
  cudaDeviceSynchronize();
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
