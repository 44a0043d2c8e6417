// Microbenchmark: what a B200 sustains for WRITE-heavy streams, by store flavour.
// The cube extract writes 8 bytes for every byte it reads from DRAM and the dense 24-channel extract is
// the one stage of the reference's dataflow that sits at 0.6 of the copy peak; torch's memset reaches
// 3.9 TB/s where a copy moves 6.5 TB/s.  This asks whether that is the memory system or the store path.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o writebw writebw.cu ; run: ./writebw
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int MODE> __device__ __forceinline__ void st4(float4* p, float4 v) {
  if (MODE == 0) *p = v;
  if (MODE == 1) asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  if (MODE == 2) asm volatile("st.global.cg.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  if (MODE == 3) asm volatile("st.global.wt.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
  if (MODE == 4) {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
  }
  if (MODE == 5) {
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
  }
}

// grid-stride float4 fill: consecutive CTAs write consecutive 4 KB runs
template <int MODE>
__global__ void __launch_bounds__(256) fill_v4(float4* __restrict__ out, size_t n4) {
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) st4<MODE>(out + i, v);
}

// scalar 4-byte stores, a warp = one 128-byte run (what the extract kernel issues)
template <int CS>
__global__ void __launch_bounds__(256) fill_s(float* __restrict__ out, size_t n) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (size_t)gridDim.x * 256) {
    if (CS) asm volatile("st.global.cs.f32 [%0], %1;" ::"l"(out + i), "f"(1.f) : "memory");
    else out[i] = 1.f;
  }
}

// each CTA owns a contiguous span (CTA-blocked instead of grid-strided)
__global__ void __launch_bounds__(256) fill_blocked(float4* __restrict__ out, size_t n4) {
  const size_t per = (n4 + gridDim.x - 1) / gridDim.x;
  const size_t b = per * blockIdx.x, e = min(n4, b + per);
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  for (size_t i = b + threadIdx.x; i < e; i += 256) out[i] = v;
}

// TMA bulk stores: shared -> global, CHUNK bytes per instruction, issued by one thread per CTA
template <int CHUNK, int DEPTH>
__global__ void __launch_bounds__(128) fill_bulk(char* __restrict__ out, size_t nbytes) {
  extern __shared__ __align__(128) char sm[];
  for (int i = threadIdx.x; i < CHUNK / 16; i += blockDim.x) reinterpret_cast<float4*>(sm)[i] = make_float4(1.f, 2.f, 3.f, 4.f);
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(sm);
    const size_t chunks = nbytes / CHUNK;
    int inflight = 0;
    for (size_t c = blockIdx.x; c < chunks; c += gridDim.x) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + c * CHUNK), "r"(s), "r"(CHUNK) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      if (++inflight >= DEPTH) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(DEPTH - 1) : "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}

// float4 copy (the "copy peak" shape) and a 1-read : 8-write mix (the extract's ratio)
__global__ void __launch_bounds__(256) copy_v4(const float4* __restrict__ in, float4* __restrict__ out, size_t n4) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) out[i] = __ldg(in + i);
}
__global__ void __launch_bounds__(256) read_v4(const float4* __restrict__ in, float* __restrict__ out, size_t n4) {
  float acc = 0.f;
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) {
    const float4 v = __ldg(in + i);
    acc += v.x + v.y + v.z + v.w;
  }
  if (acc == 123.25f) out[0] = acc;
}
__global__ void __launch_bounds__(256) mix_1r8w(const float4* __restrict__ in, float4* __restrict__ out, size_t n4_in) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4_in; i += (size_t)gridDim.x * 256) {
    const float4 v = __ldg(in + i);
#pragma unroll
    for (int r = 0; r < 8; ++r) out[(size_t)r * n4_in + i] = v;
  }
}


// the extract's output pattern: a CTA writes 32 pieces of CHUNK bytes 16 KB apart (one piece in each of 32
// a-planes of a 64^3 float cube = 512 KB for half a cube); CTAs that follow each other fill neighbouring
// pieces.  CHUNK = 1 KB is what extract_tma_kernel<4> does, 16 KB would be whole planes.
__global__ void __launch_bounds__(256) fill_planes(float4* __restrict__ out, int chunk_bytes) {
  const int pieces = 16384 / chunk_bytes;                      // pieces per plane
  const size_t half = blockIdx.x / pieces, piece = blockIdx.x % pieces;
  float4* base = out + (half * 524288 + piece * chunk_bytes) / 16;
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  const int n4 = chunk_bytes / 16;
  for (int e = threadIdx.x; e < 32 * n4; e += 256) {
    const int a = e / n4, i = e - a * n4;
    base[(size_t)a * 1024 + i] = v;
  }
}
// same bytes, but a CTA writes its 32 pieces as ONE contiguous run (32 * CHUNK bytes)
__global__ void __launch_bounds__(256) fill_runs(float4* __restrict__ out, int run_bytes) {
  float4* base = out + (size_t)blockIdx.x * run_bytes / 16;
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  for (int e = threadIdx.x; e < run_bytes / 16; e += 256) base[e] = v;
}


// L2-resident traffic: the same `span` bytes (well below the 126 MB L2) read / copied `reps` times in one launch
__global__ void __launch_bounds__(256) l2_read(const float4* __restrict__ in, float* __restrict__ out, size_t n4, int reps) {
  float acc = 0.f;
  for (int r = 0; r < reps; ++r)
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) {
      float4 v;
      asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(in + i));
      acc += v.x + v.y + v.z + v.w;
    }
  if (acc == 123.25f) out[0] = acc;
}
__global__ void __launch_bounds__(256) l2_copy(const float4* __restrict__ in, float4* __restrict__ out, size_t n4, int reps) {
  for (int r = 0; r < reps; ++r)
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) {
      float4 v;
      asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(in + i));
      out[i] = v;
    }
}
__global__ void __launch_bounds__(256) l2_write(float4* __restrict__ out, size_t n4, int reps) {
  const float4 v = make_float4(1.f, 2.f, 3.f, 4.f);
  for (int r = 0; r < reps; ++r)
    for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4; i += (size_t)gridDim.x * 256) out[i] = v;
}


// 1 read : 8 writes again, but the reads come from a window of `win4` float4 that stays in L2: DRAM sees writes only
__global__ void __launch_bounds__(256) mix_l2r8w(const float4* __restrict__ in, float4* __restrict__ out, size_t n4_in, size_t win4) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4_in; i += (size_t)gridDim.x * 256) {
    const float4 v = __ldg(in + (i % win4));
#pragma unroll
    for (int r = 0; r < 8; ++r) out[(size_t)r * n4_in + i] = v;
  }
}
// 1 read : 8 writes with the 8 copies written next to each other in time AND space (contiguous 8x expansion)
__global__ void __launch_bounds__(256) mix_1r8w_near(const float4* __restrict__ in, float4* __restrict__ out, size_t n4_in) {
  for (size_t i = (size_t)blockIdx.x * 256 + threadIdx.x; i < n4_in; i += (size_t)gridDim.x * 256) {
    const float4 v = __ldg(in + i);
    const size_t blk = i / 256, t = i % 256;     // a CTA iteration expands 4 KB into 8 consecutive 4 KB runs
#pragma unroll
    for (int r = 0; r < 8; ++r) out[(blk * 8 + r) * 256 + t] = v;
  }
}

template <typename F>
static float timeit(F launch, int reps = 10) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  launch();
  launch();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(a);
    launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    best = ms < best ? ms : best;
  }
  return best;
}

int main() {
  const size_t nbytes = (size_t)4 << 30;   // 4 GiB written per launch (>> L2)
  char* buf;
  char* src;
  CK(cudaMalloc(&buf, nbytes));
  CK(cudaMalloc(&src, nbytes));
  CK(cudaMemset(src, 1, nbytes));
  const size_t n4 = nbytes / 16;
  auto report = [&](const char* name, float ms, double bytes) { printf("%-44s %8.3f ms  %8.1f GB/s\n", name, ms, bytes / ms / 1e6); };
  report("cudaMemsetAsync", timeit([&] { cudaMemsetAsync(buf, 0, nbytes); }), (double)nbytes);
  for (int g : {148 * 2, 148 * 8, 148 * 32, 148 * 128}) {
    char nm[64];
    snprintf(nm, sizeof nm, "st.v4 default, grid %d", g);
    report(nm, timeit([&] { fill_v4<0><<<g, 256>>>((float4*)buf, n4); }), (double)nbytes);
  }
  const int G = 148 * 16;
  report("st.v4 .cs", timeit([&] { fill_v4<1><<<G, 256>>>((float4*)buf, n4); }), (double)nbytes);
  report("st.v4 .cg", timeit([&] { fill_v4<2><<<G, 256>>>((float4*)buf, n4); }), (double)nbytes);
  report("st.v4 .wt", timeit([&] { fill_v4<3><<<G, 256>>>((float4*)buf, n4); }), (double)nbytes);
  report("st.v4 L2 evict_first policy", timeit([&] { fill_v4<4><<<G, 256>>>((float4*)buf, n4); }), (double)nbytes);
  report("st.v4 L2 evict_last policy", timeit([&] { fill_v4<5><<<G, 256>>>((float4*)buf, n4); }), (double)nbytes);
  report("st.f32 default (warp = 128 B)", timeit([&] { fill_s<0><<<G, 256>>>((float*)buf, nbytes / 4); }), (double)nbytes);
  report("st.f32 .cs (warp = 128 B)", timeit([&] { fill_s<1><<<G, 256>>>((float*)buf, nbytes / 4); }), (double)nbytes);
  report("st.v4 CTA-blocked spans, grid 148*8", timeit([&] { fill_blocked<<<148 * 8, 256>>>((float4*)buf, n4); }), (double)nbytes);
  CK(cudaFuncSetAttribute(fill_bulk<32768, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
  CK(cudaFuncSetAttribute(fill_bulk<16384, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384));
  report("TMA bulk store 4 KB x4 deep, 148*8 CTAs", timeit([&] { fill_bulk<4096, 4><<<148 * 8, 128, 4096>>>(buf, nbytes); }), (double)nbytes);
  report("TMA bulk store 16 KB x4 deep, 148*4 CTAs", timeit([&] { fill_bulk<16384, 4><<<148 * 4, 128, 16384>>>(buf, nbytes); }), (double)nbytes);
  report("TMA bulk store 32 KB x4 deep, 148*2 CTAs", timeit([&] { fill_bulk<32768, 4><<<148 * 2, 128, 32768>>>(buf, nbytes); }), (double)nbytes);
  report("TMA bulk store 16 KB x8 deep, 148*8 CTAs", timeit([&] { fill_bulk<16384, 8><<<148 * 8, 128, 16384>>>(buf, nbytes); }), (double)nbytes);
  for (int c : {128, 256, 512, 1024, 2048, 4096, 16384}) {
    char nm[80];
    snprintf(nm, sizeof nm, "plane-strided pieces of %d B x 32 planes / CTA", c);
    const unsigned g = (unsigned)(nbytes / (32ull * c));
    report(nm, timeit([&] { fill_planes<<<g, 256>>>((float4*)buf, c); }), (double)nbytes);
  }
  for (int c : {4096, 32768, 524288}) {
    char nm[80];
    snprintf(nm, sizeof nm, "contiguous run of %d B / CTA", c);
    const unsigned g = (unsigned)(nbytes / (size_t)c);
    report(nm, timeit([&] { fill_runs<<<g, 256>>>((float4*)buf, c); }), (double)nbytes);
  }
  {
    const size_t span = (size_t)24 << 20;   // 24 MiB read (+ 24 MiB written): L2-resident
    const int reps = 64;
    report("L2-resident read 24 MiB x64", timeit([&] { l2_read<<<G, 256>>>((const float4*)src, (float*)buf, span / 16, reps); }), (double)span * reps);
    report("L2-resident write 24 MiB x64", timeit([&] { l2_write<<<G, 256>>>((float4*)buf, span / 16, reps); }), (double)span * reps);
    report("L2-resident copy 24+24 MiB x64 (r+w bytes)", timeit([&] { l2_copy<<<G, 256>>>((const float4*)src, (float4*)buf, span / 16, reps); }), (double)span * reps * 2);
  }
  report("read only (ld.v4)", timeit([&] { read_v4<<<G, 256>>>((const float4*)src, (float*)buf, n4); }), (double)nbytes);
  report("copy 2+2 GiB (ld.v4 -> st.v4)", timeit([&] { copy_v4<<<G, 256>>>((const float4*)src, (float4*)buf, n4 / 2); }), (double)nbytes);
  report("mix 1 read : 8 writes (0.44 + 3.56 GiB)", timeit([&] { mix_1r8w<<<G, 256>>>((const float4*)src, (float4*)buf, n4 / 9); }), (double)(n4 / 9) * 16 * 9);
  report("mix 1 L2-resident read : 8 writes (3.56 GiB written)", timeit([&] { mix_l2r8w<<<G, 256>>>((const float4*)src, (float4*)buf, n4 / 9, (size_t)(16 << 20) / 16); }), (double)(n4 / 9) * 16 * 8);
  report("mix 1 read : 8 writes, copies adjacent (r+w bytes)", timeit([&] { mix_1r8w_near<<<G, 256>>>((const float4*)src, (float4*)buf, n4 / 9); }), (double)(n4 / 9) * 16 * 9);
  CK(cudaDeviceSynchronize());
  CK(cudaGetLastError());
  return 0;
}
