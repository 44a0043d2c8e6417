// R5/R6: fused cube extraction into the model's input batch.
//
// Replaces GridCreator.transpose + create_grids_from_mrc (utils/create_grids.py:67-176,
// reference root), the five training twins (scripts_for_training_data/create_grids_for_*.py)
// and CryoEMTestDataset.__getitem__ (dataset/dataset.py:194-224): instead of np.pad +
// one .npz per cube per channel + 25 np.load per cube, every W^3 window (W = grid_size +
// 2*padding) is cut straight from the resident (nz,ny,nx) volumes into the
// [B, C, W, W, W] tensor the model consumes, zero-filled outside the map.
//
// The reference cuts cubes from an axis-permuted view (D6): cube axis m walks memory
// axis perm[m].  For the standard MRC axis order perm = (2,1,0), i.e. the cube's slowest
// axis is the memory-contiguous one, so the copy is a tiled transpose through shared
// memory (32x33 tiles: coalesced 128-byte reads along x, coalesced 128-byte writes
// along the cube's fastest axis).  Pure index work: bit-exact.
#include <math_constants.h>

#include "common.cuh"

namespace mica {

struct ExtractParams {
  const float* vol;
  int64_t chan_stride;
  int T[3];            // cube-space global dims: T[m] = memdims[perm[m]]
  int64_t sstride[3];  // memory stride walked by cube axis m
  int slab_axis;       // cube axis that walks memory axis 0 (the slab axis)
  int z0, nzl;         // slab: memory planes [z0, z0 + nzl)
  int W, pad;
  const int32_t* ijk;
  float* out;
  int64_t out_cube_stride;
  int32_t* nonzero;
  float* cube_max;
};

__device__ __forceinline__ void atomic_max_f32(float* addr, float v) {
  if (v >= 0.f)
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else
    atomicMin(reinterpret_cast<unsigned*>(addr), __float_as_uint(v));
}

__device__ __forceinline__ void block_flags(const ExtractParams& P, int cube, bool any_nz, float vmax) {
  if (P.nonzero) {
    if (__syncthreads_or(any_nz) && threadIdx.x == 0 && threadIdx.y == 0) atomicOr(&P.nonzero[cube], 1);
  }
  if (P.cube_max && blockIdx.y == 0) {
    for (int o = 16; o > 0; o >>= 1) vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    if (threadIdx.x == 0) atomic_max_f32(&P.cube_max[cube], vmax);
  }
}

// memory offset of cube-space position p (relative to the local slab), or -1 outside the map
__device__ __forceinline__ bool in_map(const ExtractParams& P, int p0, int p1, int p2) {
  return (unsigned)p0 < (unsigned)P.T[0] && (unsigned)p1 < (unsigned)P.T[1] && (unsigned)p2 < (unsigned)P.T[2];
}

// ---- cube's fastest axis (2) is the memory-contiguous one: row copy.
// grid = (W [u0], C, B), block = (32, 8)
__global__ void __launch_bounds__(256)
extract_rows_kernel(ExtractParams P) {
  const int u0 = blockIdx.x, c = blockIdx.y, b = blockIdx.z;
  const int W = P.W;
  const int i0 = P.ijk[3 * b + 0] - P.pad, j0 = P.ijk[3 * b + 1] - P.pad, k0 = P.ijk[3 * b + 2] - P.pad;
  const float* src = P.vol + (int64_t)c * P.chan_stride;
  float* dst = P.out + (int64_t)b * P.out_cube_stride + (int64_t)c * W * W * W + (int64_t)u0 * W * W;
  const int p0 = i0 + u0;
  const int zsub[3] = {P.slab_axis == 0 ? P.z0 : 0, P.slab_axis == 1 ? P.z0 : 0, P.slab_axis == 2 ? P.z0 : 0};
  bool any_nz = false;
  float vmax = -CUDART_INF_F;
  for (int u1 = threadIdx.y; u1 < W; u1 += blockDim.y) {
    const int p1 = j0 + u1;
    for (int u2 = threadIdx.x; u2 < W; u2 += 32) {
      const int p2 = k0 + u2;
      float v = 0.f;
      if (in_map(P, p0, p1, p2)) {
        int q0 = p0 - zsub[0], q1 = p1 - zsub[1], q2 = p2 - zsub[2];
        int qs = P.slab_axis == 0 ? q0 : (P.slab_axis == 1 ? q1 : q2);
        if ((unsigned)qs < (unsigned)P.nzl) v = src[q0 * P.sstride[0] + q1 * P.sstride[1] + q2 * P.sstride[2]];
      }
      any_nz |= (v != 0.f);
      vmax = fmaxf(vmax, v);
      st_stream(dst + (int64_t)u1 * W + u2, v);
    }
  }
  block_flags(P, b, any_nz, vmax);
}

// ---- cube axis MC (0 or 1) is memory-contiguous: tiled transpose MC <-> 2.
// grid = (W [u_o, the other slow axis], C, B), block = (32, 8)
template <int MC>
__global__ void __launch_bounds__(256)
extract_transpose_kernel(ExtractParams P) {
  __shared__ float tile[32][33];
  constexpr int OA = 1 - MC;
  const int uo = blockIdx.x, c = blockIdx.y, b = blockIdx.z;
  const int W = P.W;
  int org[3] = {P.ijk[3 * b + 0] - P.pad, P.ijk[3 * b + 1] - P.pad, P.ijk[3 * b + 2] - P.pad};
  const float* src = P.vol + (int64_t)c * P.chan_stride;
  float* dst = P.out + (int64_t)b * P.out_cube_stride + (int64_t)c * W * W * W;
  const int64_t out_mc = (MC == 0) ? (int64_t)W * W : W;   // output stride of cube axis MC
  const int64_t out_oa = (MC == 0) ? W : (int64_t)W * W;   // output stride of the other slow axis
  const int po = org[OA] + uo;
  const bool ok_o = (unsigned)po < (unsigned)P.T[OA];
  const int zs_o = (P.slab_axis == OA) ? P.z0 : 0, zs_c = (P.slab_axis == MC) ? P.z0 : 0,
            zs_2 = (P.slab_axis == 2) ? P.z0 : 0;
  const bool slab_o = (P.slab_axis != OA) || ((unsigned)(po - P.z0) < (unsigned)P.nzl);
  const int64_t base_o = (int64_t)(po - zs_o) * P.sstride[OA];
  const int tx = threadIdx.x, ty = threadIdx.y;
  bool any_nz = false;
  float vmax = -CUDART_INF_F;
  for (int t2 = 0; t2 < W; t2 += 32) {
    for (int tc = 0; tc < W; tc += 32) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int u2 = t2 + ty + 8 * r, uc = tc + tx;
        const int p2 = org[2] + u2, pc = org[MC] + uc;
        float v = 0.f;
        if (ok_o && slab_o && u2 < W && uc < W && (unsigned)p2 < (unsigned)P.T[2] && (unsigned)pc < (unsigned)P.T[MC]) {
          const int q2 = p2 - zs_2, qc = pc - zs_c;
          const bool slab_ok = (P.slab_axis == 2) ? ((unsigned)q2 < (unsigned)P.nzl)
                                                  : (P.slab_axis == MC ? ((unsigned)qc < (unsigned)P.nzl) : true);
          if (slab_ok) v = __ldg(src + base_o + (int64_t)q2 * P.sstride[2] + (int64_t)qc * P.sstride[MC]);
        }
        any_nz |= (v != 0.f);
        vmax = fmaxf(vmax, v);
        tile[ty + 8 * r][tx] = v;
      }
      __syncthreads();
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int uc = tc + ty + 8 * r, u2 = t2 + tx;
        if (uc < W && u2 < W) st_stream(dst + uc * out_mc + uo * out_oa + u2, tile[tx][ty + 8 * r]);
      }
      __syncthreads();
    }
  }
  block_flags(P, b, any_nz, vmax);
}

__global__ void init_flags_kernel(int32_t* nonzero, float* cube_max, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (nonzero) nonzero[i] = 0;
  if (cube_max) cube_max[i] = -CUDART_INF_F;
}

}  // namespace mica

using namespace mica;

extern "C" int mica_extract_cubes(const float* vol, int64_t chan_stride, int n_channels,
                                  int nz, int ny, int nx, int z0, int nz_local, const int perm[3],
                                  int grid_size, int padding, const int32_t* ijk, int n_cubes,
                                  float* out, int64_t out_cube_stride, int32_t* nonzero, float* cube_max,
                                  mica_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MICA_REQUIRE(vol && perm && out && (ijk || n_cubes == 0), "null pointer");
  MICA_REQUIRE(n_channels > 0 && n_channels <= 65535, "bad channel count %d", n_channels);
  MICA_REQUIRE(nz > 0 && ny > 0 && nx > 0, "empty volume");
  MICA_REQUIRE(z0 >= 0 && nz_local > 0 && z0 + nz_local <= nz, "bad slab");
  MICA_REQUIRE(grid_size > 0 && padding >= 0, "bad grid_size/padding");
  int seen = 0;
  for (int m = 0; m < 3; ++m) {
    MICA_REQUIRE(perm[m] >= 0 && perm[m] < 3, "perm must be a permutation of 0,1,2");
    seen |= 1 << perm[m];
  }
  MICA_REQUIRE(seen == 7, "perm must be a permutation of 0,1,2");
  const int W = grid_size + 2 * padding;
  MICA_REQUIRE(W <= 65535, "window too large");
  if (n_cubes <= 0) return MICA_OK;

  ExtractParams P;
  const int memdims[3] = {nz, ny, nx};
  const int64_t memstride[3] = {(int64_t)ny * nx, nx, 1};
  P.vol = vol;
  P.chan_stride = chan_stride;
  P.slab_axis = 0;
  for (int m = 0; m < 3; ++m) {
    P.T[m] = memdims[perm[m]];
    P.sstride[m] = memstride[perm[m]];
    if (perm[m] == 0) P.slab_axis = m;
  }
  P.z0 = z0;
  P.nzl = nz_local;
  P.W = W;
  P.pad = padding;
  P.out = out;
  P.out_cube_stride = out_cube_stride;
  P.nonzero = nonzero;
  P.cube_max = cube_max;

  const int contiguous_axis = perm[0] == 2 ? 0 : (perm[1] == 2 ? 1 : 2);
  const int kMaxZ = 32768;
  for (int b0 = 0; b0 < n_cubes; b0 += kMaxZ) {
    const int nb = (n_cubes - b0 < kMaxZ) ? n_cubes - b0 : kMaxZ;
    P.ijk = ijk + 3 * (int64_t)b0;
    P.out = out + (int64_t)b0 * out_cube_stride;
    P.nonzero = nonzero ? nonzero + b0 : nullptr;
    P.cube_max = cube_max ? cube_max + b0 : nullptr;
    if (nonzero || cube_max) {
      init_flags_kernel<<<(nb + 255) / 256, 256, 0, st>>>(P.nonzero, P.cube_max, nb);
      MICA_LAUNCH_CHECK("init_flags_kernel");
    }
    dim3 grid(W, n_channels, nb), block(32, 8);
    if (contiguous_axis == 2)
      extract_rows_kernel<<<grid, block, 0, st>>>(P);
    else if (contiguous_axis == 0)
      extract_transpose_kernel<0><<<grid, block, 0, st>>>(P);
    else
      extract_transpose_kernel<1><<<grid, block, 0, st>>>(P);
    MICA_LAUNCH_CHECK("extract_cubes kernel");
  }
  return MICA_OK;
}
