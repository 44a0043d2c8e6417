// R4: 24-channel AF3 rasteriser.
//
// Replaces transform_coordinates + the per-atom Python loop at
// utils/preprocessing.py:172-178,275-298 (reference root).  Each atom marks one
// voxel in its backbone channel (CA,N,C,O) and one in its residue-type channel.
// Stores of 1.0f are idempotent, so the scatter needs no atomics and is
// order-independent (bit-exact).  Traffic is the zero fill (96 B/voxel); the
// scatter itself touches < 2 sectors per atom.
#include "common.cuh"

namespace mica {

__global__ void __launch_bounds__(256)
fill_zero_kernel(float4* __restrict__ p4, long long n4, float* __restrict__ tail, int ntail) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) p4[i] = z;
  if (blockIdx.x == 0 && threadIdx.x < ntail) tail[threadIdx.x] = 0.f;
}

int fill_zero(float* p, long long n, cudaStream_t st) {
  if (n <= 0) return MICA_OK;
  // align the head to 16 bytes with the scalar tail path (cudaMalloc'd buffers already are)
  long long head = (((uintptr_t)p & 15) == 0) ? 0 : ((16 - ((uintptr_t)p & 15)) / 4);
  if (head > n) head = n;
  if (head) {
    fill_zero_kernel<<<1, 32, 0, st>>>(nullptr, 0, p, (int)head);
    MICA_LAUNCH_CHECK("fill_zero_kernel(head)");
  }
  float* q = p + head;
  long long m = n - head, n4 = m >> 2;
  int64_t want = ceil_div64(n4 > 0 ? n4 : 1, 256 * 4);
  int grid = (int)(want < (int64_t)kNumSMs * 8 ? want : (int64_t)kNumSMs * 8);
  fill_zero_kernel<<<grid, 256, 0, st>>>((float4*)q, n4, q + n4 * 4, (int)(m & 3));
  MICA_LAUNCH_CHECK("fill_zero_kernel");
  return MICA_OK;
}

__global__ void __launch_bounds__(256)
af3_scatter_kernel(const float* __restrict__ xyz, const int8_t* __restrict__ bb_ch,
                   const int8_t* __restrict__ aa_ch, long long n_atoms, float ox, float oy, float oz,
                   int clip_x, int clip_y, int clip_z, int nz, int ny, int nx, int z0, int nz_local,
                   float* __restrict__ vol, int* __restrict__ status_oob) {
  long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_atoms) return;
  // np.round(coord - origin) in float32, round-half-even; then clip(., 0, shape - 1)
  float fx = rintf(__fsub_rn(xyz[3 * a + 0], ox));
  float fy = rintf(__fsub_rn(xyz[3 * a + 1], oy));
  float fz = rintf(__fsub_rn(xyz[3 * a + 2], oz));
  // float -> int64 as numpy astype(int); clamp in float first so the cast cannot overflow
  long long ix = (long long)fminf(fmaxf(fx, -4.0e18f), 4.0e18f);
  long long iy = (long long)fminf(fmaxf(fy, -4.0e18f), 4.0e18f);
  long long iz = (long long)fminf(fmaxf(fz, -4.0e18f), 4.0e18f);
  ix = ix < 0 ? 0 : (ix > clip_x ? clip_x : ix);
  iy = iy < 0 ? 0 : (iy > clip_y ? clip_y : iy);
  iz = iz < 0 ? 0 : (iz > clip_z ? clip_z : iz);
  const int b = bb_ch[a], r = aa_ch[a];
  if (b < 0 && r < 0) return;                 // the reference indexes the volume only when it writes
  if (ix >= nx || iy >= ny || iz >= nz) {     // numpy would raise IndexError (D7) -> encoding fails
    atomicExch(status_oob, 1);
    return;
  }
  const long long zl = iz - z0;
  if (zl < 0 || zl >= nz_local) return;       // another rank's slab
  const long long chan = (long long)nz_local * ny * nx;
  const long long off = (zl * ny + iy) * nx + ix;
  if (b >= 0) vol[b * chan + off] = 1.0f;
  if (r >= 0) vol[r * chan + off] = 1.0f;
}

}  // namespace mica

using namespace mica;

extern "C" int mica_af3_encode(const float* xyz, const int8_t* bb_ch, const int8_t* aa_ch, int64_t n_atoms,
                               float ox, float oy, float oz, int clip_x, int clip_y, int clip_z,
                               int nz, int ny, int nx, int z0, int nz_local,
                               float* vol24, int* status_oob, mica_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MICA_REQUIRE(vol24 && status_oob, "null pointer");
  MICA_REQUIRE(n_atoms == 0 || (xyz && bb_ch && aa_ch), "null atom arrays");
  MICA_REQUIRE(nz > 0 && ny > 0 && nx > 0, "empty grid");
  MICA_REQUIRE(z0 >= 0 && nz_local >= 0 && z0 + nz_local <= nz, "bad slab");
  MICA_REQUIRE(clip_x >= 0 && clip_y >= 0 && clip_z >= 0, "negative clip bound");
  MICA_CUDA(cudaMemsetAsync(status_oob, 0, sizeof(int), st));
  int rc = fill_zero(vol24, 24LL * nz_local * ny * nx, st);
  if (rc) return rc;
  if (n_atoms > 0 && nz_local > 0) {
    af3_scatter_kernel<<<(unsigned)ceil_div64(n_atoms, 256), 256, 0, st>>>(
        xyz, bb_ch, aa_ch, n_atoms, ox, oy, oz, clip_x, clip_y, clip_z, nz, ny, nx, z0, nz_local, vol24,
        status_oob);
    MICA_LAUNCH_CHECK("af3_scatter_kernel");
  }
  return MICA_OK;
}
