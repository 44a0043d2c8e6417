// Microbenchmark: how fast one SM's TMA unit turns box loads into bytes, by row length and ring depth.
// The cube extract loads [BT b][64 c][32 x] boxes (128-byte rows, SWIZZLE_128B) and four differently built
// kernels all stop at the same time per batch; this isolates the load side: a persistent CTA per SM issues box
// loads of a 3-D float tensor into a ring of DEPTH stages and does nothing with them.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmarate tmarate.cu ; run: ./tmarate
#include <cstdint>
#include <cstdio>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// volume [nz][ny][nx] floats; box = [bz][by][bx]; a CTA walks boxes b = blockIdx.x, + gridDim.x, ...
template <int DEPTH>
__global__ void __launch_bounds__(128) box_loads(const __grid_constant__ CUtensorMap tmap, int box_bytes, int nbx, int nby,
                                                 int bx, int by, int bz, long long n_boxes, int stride_boxes) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[DEPTH];
  const uint32_t ring = (smem_u32(smem_raw) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < DEPTH; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  auto issue = [&](long long b, int slot) {
    // boxes overlap like the extract's windows when stride_boxes < box edge: b enumerates (z, y, x) box origins
    const long long bxy = (long long)nbx * nby;
    const int iz = (int)(b / bxy), iy = (int)((b % bxy) / nbx), ix = (int)(b % nbx);
    const uint32_t bar = smem_u32(&full[slot]);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(box_bytes) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(ring + (uint32_t)slot * (uint32_t)box_bytes), "l"(&tmap), "r"(ix * bx), "r"(iy * by), "r"(iz * bz), "r"(bar)
        : "memory");
  };
  long long b = blockIdx.x;
  int issued = 0;
  for (; issued < DEPTH && b < n_boxes; ++issued, b += gridDim.x) issue(b, issued);
  int slot = 0;
  uint32_t phase = 0;
  for (long long done = blockIdx.x; done < n_boxes; done += gridDim.x) {
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(&full[slot])), "r"(phase) : "memory");
    if (b < n_boxes) {
      issue(b, slot);
      b += gridDim.x;
    }
    if (++slot == DEPTH) {
      slot = 0;
      phase ^= 1u;
    }
  }
}

template <int DEPTH>
static float run(const CUtensorMap& tmap, int box_bytes, int nbx, int nby, int nbz, int bx, int by, int bz, int ctas_per_sm) {
  const size_t smem = (size_t)DEPTH * box_bytes + 1024;
  cudaFuncSetAttribute(box_loads<DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const long long n_boxes = (long long)nbx * nby * nbz;
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  float best = 1e30f;
  for (int r = 0; r < 6; ++r) {
    cudaEventRecord(a);
    box_loads<DEPTH><<<148 * ctas_per_sm, 128, smem>>>(tmap, box_bytes, nbx, nby, bx, by, bz, n_boxes, 0);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (r > 0 && ms < best) best = ms;
  }
  return best;
}

int main() {
  EncodeFn enc = nullptr;
  cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
  const int nx = 1024, ny = 1024, nz = 512;                 // 2 GiB of floats
  float* vol;
  CK(cudaMalloc(&vol, (size_t)nx * ny * nz * 4));
  CK(cudaMemset(vol, 0, (size_t)nx * ny * nz * 4));
  struct Case { const char* name; int bx, by, bz; CUtensorMapSwizzle sw; };
  const Case cases[] = {
      {"rows of 128 B (32 x), box 32x64x4, SWIZZLE_128B (the extract's box)", 32, 64, 4, CU_TENSOR_MAP_SWIZZLE_128B},
      {"rows of 128 B (32 x), box 32x64x4, no swizzle", 32, 64, 4, CU_TENSOR_MAP_SWIZZLE_NONE},
      {"rows of 256 B (64 x), box 64x64x2, no swizzle", 64, 64, 2, CU_TENSOR_MAP_SWIZZLE_NONE},
      {"rows of 512 B (128 x), box 128x64x1, no swizzle", 128, 64, 1, CU_TENSOR_MAP_SWIZZLE_NONE},
      {"rows of 1 KB (256 x), box 256x32x1, no swizzle", 256, 32, 1, CU_TENSOR_MAP_SWIZZLE_NONE},
      {"rows of 64 B (16 x), box 16x64x8, SWIZZLE_64B", 16, 64, 8, CU_TENSOR_MAP_SWIZZLE_64B},
  };
  for (const Case& c : cases) {
    CUtensorMap tmap;
    cuuint64_t gdim[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)nz};
    cuuint64_t gstr[2] = {(cuuint64_t)nx * 4, (cuuint64_t)nx * ny * 4};
    cuuint32_t box[3] = {(cuuint32_t)c.bx, (cuuint32_t)c.by, (cuuint32_t)c.bz};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, vol, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, c.sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      printf("%-72s encode failed (%d)\n", c.name, (int)r);
      continue;
    }
    const int box_bytes = c.bx * c.by * c.bz * 4;          // 32 KB in every case
    const int nbx = nx / c.bx, nby = ny / c.by, nbz = nz / c.bz;
    const double bytes = (double)nx * ny * nz * 4;
    printf("%s\n", c.name);
    printf("   1 CTA/SM: depth 2 %7.1f GB/s   depth 4 %7.1f GB/s   depth 6 %7.1f GB/s\n",
           bytes / run<2>(tmap, box_bytes, nbx, nby, nbz, c.bx, c.by, c.bz, 1) / 1e6,
           bytes / run<4>(tmap, box_bytes, nbx, nby, nbz, c.bx, c.by, c.bz, 1) / 1e6,
           bytes / run<6>(tmap, box_bytes, nbx, nby, nbz, c.bx, c.by, c.bz, 1) / 1e6);
    printf("   2 CTA/SM: depth 3 %7.1f GB/s   4 CTA/SM: depth 1 %7.1f GB/s\n",
           bytes / run<3>(tmap, box_bytes, nbx, nby, nbz, c.bx, c.by, c.bz, 2) / 1e6,
           bytes / run<1>(tmap, box_bytes, nbx, nby, nbz, c.bx, c.by, c.bz, 4) / 1e6);
  }
  CK(cudaDeviceSynchronize());
  CK(cudaGetLastError());
  return 0;
}
