// SURVEY 8(f) N3 + N4: label-mask rasterisers of the training-data builders and the docking masks.
//
// N3 replaces the Python dictionaries + loops of
//   scripts_for_training_data/create_backbone_mask.py:136-172      (3 = backbone atom, 2 = other atom,
//   scripts_for_training_data/create_carbon_alpha_mask.py:136-173   1 = 26-neighbour of an atom voxel)
//   scripts_for_training_data/create_amino_acid_mask.py:151-177    (residue label on the 26 neighbours of
//                                                                   every C-alpha, lowest label wins,
//                                                                   C-alpha voxels zeroed in file order)
// Both are ORDER-DEPENDENT in the reference (last writer wins / sequential zeroing).  The kernels use
// order-free statements that give the same volume bit for bit:
//   * class mask: per voxel the atom with the highest file index wins -> atomicMax on (index << 2 | class);
//   * amino-acid mask: per voxel v, a_all = min label over neighbouring C-alphas, T = last C-alpha ON v,
//     a_0 = min label over neighbouring C-alphas that come before T; v ends as a_all if T does not exist
//     or a_all < a_0 (a later C-alpha lowered the minimum after the zeroing), else 0.
//     (oracle/masks_oracle.py::amino_acid_mask_closed_form checks this against the sequential loop.)
// N4 replaces utils/dock_in_map.py:269 (contour threshold) and :330-352 (zero within `radius` of the
// selected atoms): instead of a full Euclidean distance transform of the map, every seed voxel zeroes the
// voxels of its own ball, the distance evaluated exactly as SciPy does (float64, (dz*s0)^2 + (dy*s1)^2
// first, then + (dx*s2)^2, sqrt, <= radius).
#include "common.cuh"

namespace mica {

// float32 subtract, round half to even, clip against the caller's (mis-ordered) bounds
__device__ __forceinline__ bool atom_voxel(const float* __restrict__ xyz, long long a, float ox, float oy, float oz,
                                           int clip_x, int clip_y, int clip_z, int nz, int ny, int nx, int& ix,
                                           int& iy, int& iz) {
  const float fx = rintf(__fsub_rn(xyz[3 * a + 0], ox));
  const float fy = rintf(__fsub_rn(xyz[3 * a + 1], oy));
  const float fz = rintf(__fsub_rn(xyz[3 * a + 2], oz));
  long long lx = (long long)fminf(fmaxf(fx, -4.0e18f), 4.0e18f);
  long long ly = (long long)fminf(fmaxf(fy, -4.0e18f), 4.0e18f);
  long long lz = (long long)fminf(fmaxf(fz, -4.0e18f), 4.0e18f);
  lx = lx < 0 ? 0 : (lx > clip_x ? clip_x : lx);
  ly = ly < 0 ? 0 : (ly > clip_y ? clip_y : ly);
  lz = lz < 0 ? 0 : (lz > clip_z ? clip_z : lz);
  ix = (int)lx;
  iy = (int)ly;
  iz = (int)lz;
  return lx < nx && ly < ny && lz < nz;  // false: numpy raises IndexError on `mask[pos] = ...`
}

// pass 1: the last atom (highest file index) on a voxel decides its class
__global__ void __launch_bounds__(256)
class_mask_atoms_kernel(const float* __restrict__ xyz, const uint8_t* __restrict__ is_class, long long n_atoms,
                        float ox, float oy, float oz, int clip_x, int clip_y, int clip_z, int nz, int ny, int nx,
                        int* __restrict__ mask, int* __restrict__ status_oob) {
  const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_atoms) return;
  int ix, iy, iz;
  if (!atom_voxel(xyz, a, ox, oy, oz, clip_x, clip_y, clip_z, nz, ny, nx, ix, iy, iz)) {
    atomicExch(status_oob, 1);
    return;
  }
  const int key = (int)((a + 1) << 2) | (is_class[a] ? 3 : 2);
  atomicMax(&mask[((long long)iz * ny + iy) * nx + ix], key);
}

// pass 2: keys (>= 4) become their class, empty 26-neighbours of atom voxels become 1
__global__ void __launch_bounds__(256)
class_mask_finish_kernel(const float* __restrict__ xyz, long long n_atoms, float ox, float oy, float oz,
                         int clip_x, int clip_y, int clip_z, int nz, int ny, int nx, int* mask) {
  const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= n_atoms) return;
  int ix, iy, iz;
  if (!atom_voxel(xyz, a, ox, oy, oz, clip_x, clip_y, clip_z, nz, ny, nx, ix, iy, iz)) return;
  for (int dz = -1; dz <= 1; ++dz) {
    const int z = iz + dz;
    if (z < 0 || z >= nz) continue;
    for (int dy = -1; dy <= 1; ++dy) {
      const int y = iy + dy;
      if (y < 0 || y >= ny) continue;
      for (int dx = -1; dx <= 1; ++dx) {
        const int x = ix + dx;
        if (x < 0 || x >= nx) continue;
        int* p = &mask[((long long)z * ny + y) * nx + x];
        if (dz == 0 && dy == 0 && dx == 0) {
          const int v = atomicAdd(p, 0);
          if (v >= 4) atomicCAS(p, v, v & 3);  // same result whichever atom of the voxel gets here first
        } else {
          atomicCAS(p, 0, 1);  // only an untouched voxel: atom voxels hold a key (>= 4) or a class (2, 3)
        }
      }
    }
  }
}

constexpr int kAaInf = 0x7f7f7f7f;  // cudaMemset(0x7f) pattern = "no label yet"

__global__ void __launch_bounds__(256)
aa_mask_pass1_kernel(const float* __restrict__ xyz, const int32_t* __restrict__ label, long long n_ca, float ox,
                     float oy, float oz, int clip_x, int clip_y, int clip_z, int nz, int ny, int nx,
                     int* __restrict__ a_all, int* __restrict__ last_on, int* __restrict__ status_oob) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_ca) return;
  int ix, iy, iz;
  if (!atom_voxel(xyz, t, ox, oy, oz, clip_x, clip_y, clip_z, nz, ny, nx, ix, iy, iz)) {
    // the reference only indexes the mask at the C-alpha voxel itself (`mask[ca_pos] = 0`), after the
    // in-bounds neighbours were assigned; the IndexError aborts generate_mask all the same
    atomicExch(status_oob, 1);
    return;
  }
  const int lab = label[t];
  atomicMax(&last_on[((long long)iz * ny + iy) * nx + ix], (int)(t + 1));
  for (int dz = -1; dz <= 1; ++dz) {
    const int z = iz + dz;
    if (z < 0 || z >= nz) continue;
    for (int dy = -1; dy <= 1; ++dy) {
      const int y = iy + dy;
      if (y < 0 || y >= ny) continue;
      for (int dx = -1; dx <= 1; ++dx) {
        const int x = ix + dx;
        if (x < 0 || x >= nx || (dz == 0 && dy == 0 && dx == 0)) continue;
        atomicMin(&a_all[((long long)z * ny + y) * nx + x], lab);
      }
    }
  }
}

__global__ void __launch_bounds__(256)
aa_mask_pass2_kernel(const float* __restrict__ xyz, const int32_t* __restrict__ label, long long n_ca, float ox,
                     float oy, float oz, int clip_x, int clip_y, int clip_z, int nz, int ny, int nx,
                     const int* __restrict__ last_on, int* __restrict__ a_0) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_ca) return;
  int ix, iy, iz;
  if (!atom_voxel(xyz, t, ox, oy, oz, clip_x, clip_y, clip_z, nz, ny, nx, ix, iy, iz)) return;
  const int lab = label[t];
  for (int dz = -1; dz <= 1; ++dz) {
    const int z = iz + dz;
    if (z < 0 || z >= nz) continue;
    for (int dy = -1; dy <= 1; ++dy) {
      const int y = iy + dy;
      if (y < 0 || y >= ny) continue;
      for (int dx = -1; dx <= 1; ++dx) {
        const int x = ix + dx;
        if (x < 0 || x >= nx || (dz == 0 && dy == 0 && dx == 0)) continue;
        const long long v = ((long long)z * ny + y) * nx + x;
        const int last = last_on[v];
        if (last != 0 && (int)(t + 1) < last) atomicMin(&a_0[v], lab);
      }
    }
  }
}

__global__ void __launch_bounds__(256)
aa_mask_pass3_kernel(const float* __restrict__ xyz, long long n_ca, float ox, float oy, float oz, int clip_x,
                     int clip_y, int clip_z, int nz, int ny, int nx, const int* __restrict__ a_all,
                     const int* __restrict__ last_on, const int* __restrict__ a_0, int* __restrict__ mask) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_ca) return;
  int ix, iy, iz;
  if (!atom_voxel(xyz, t, ox, oy, oz, clip_x, clip_y, clip_z, nz, ny, nx, ix, iy, iz)) return;
  for (int dz = -1; dz <= 1; ++dz) {
    const int z = iz + dz;
    if (z < 0 || z >= nz) continue;
    for (int dy = -1; dy <= 1; ++dy) {
      const int y = iy + dy;
      if (y < 0 || y >= ny) continue;
      for (int dx = -1; dx <= 1; ++dx) {
        const int x = ix + dx;
        if (x < 0 || x >= nx) continue;
        const long long v = ((long long)z * ny + y) * nx + x;
        const int all = a_all[v];
        int out = 0;
        if (all != kAaInf) out = (last_on[v] == 0 || all < a_0[v]) ? all : 0;
        mask[v] = out;  // every writer of a voxel stores the same value
      }
    }
  }
}

// ------------------------------------------------------------------------------------ N4
__global__ void __launch_bounds__(256)
contour_threshold_kernel(const float* __restrict__ in, float* __restrict__ out, long long n, float level) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = in[i];
    out[i] = (v < level) ? 0.f : v;  // np.where(data < level, 0, data): NaN stays
  }
}

// one warp per selected atom; lanes sweep the bounding box of its ball
__global__ void __launch_bounds__(256)
zero_around_atoms_kernel(const float* __restrict__ xyz, long long n_atoms, float ox, float oy, float oz, float vx,
                         float vy, float vz, double s0, double s1, double s2, double radius, int rz, int ry, int rx,
                         int nz, int ny, int nx, float* map, int* __restrict__ status_oob) {
  const long long a = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (a >= n_atoms) return;
  // ((coords - origin) / voxel_size).astype(int): float32 arithmetic, truncation toward zero
  const float fx = __fdiv_rn(__fsub_rn(xyz[3 * a + 0], ox), vx);
  const float fy = __fdiv_rn(__fsub_rn(xyz[3 * a + 1], oy), vy);
  const float fz = __fdiv_rn(__fsub_rn(xyz[3 * a + 2], oz), vz);
  const long long lx = (long long)fminf(fmaxf(fx, -4.0e18f), 4.0e18f);
  const long long ly = (long long)fminf(fmaxf(fy, -4.0e18f), 4.0e18f);
  const long long lz = (long long)fminf(fmaxf(fz, -4.0e18f), 4.0e18f);
  // valid_coords: (x,y,z) >= 0 and < map.shape = (nz,ny,nx) -- compared in that (mis-matched) order
  if (lx < 0 || ly < 0 || lz < 0 || lx >= nz || ly >= ny || lz >= nx) return;
  if (lx >= nx || lz >= nz) {  // passes the test but mask[z, y, x] is out of bounds: IndexError in the reference
    if (lane == 0) atomicExch(status_oob, 1);
    return;
  }
  const int cx = (int)lx, cy = (int)ly, cz = (int)lz;
  const int bx = 2 * rx + 1, by = 2 * ry + 1, bz = 2 * rz + 1;
  const int box = bx * by * bz;
  for (int k = lane; k < box; k += 32) {
    const int dx = k % bx - rx, dy = (k / bx) % by - ry, dz = k / (bx * by) - rz;
    const int x = cx + dx, y = cy + dy, z = cz + dz;
    if (x < 0 || x >= nx || y < 0 || y >= ny || z < 0 || z >= nz) continue;
    const double tz = __dmul_rn((double)dz, s0), ty = __dmul_rn((double)dy, s1), tx = __dmul_rn((double)dx, s2);
    const double d2 = __dadd_rn(__dadd_rn(__dmul_rn(tz, tz), __dmul_rn(ty, ty)), __dmul_rn(tx, tx));
    if (__dsqrt_rn(d2) <= radius) map[((long long)z * ny + y) * nx + x] = 0.f;
  }
}

static inline unsigned blocks_for(long long n, int threads) { return (unsigned)ceil_div64(n > 0 ? n : 1, threads); }

}  // namespace mica

using namespace mica;

/* N3 class masks.  xyz: device float32 [A,3] (x,y,z) of ALL atoms in file order; is_class: device uint8 [A]
 * (1 = backbone atom N/CA/C/O for the backbone mask, 1 = CA for the C-alpha mask).  mask: device int32
 * [nz,ny,nx], fully written.  status_oob: device int, 1 when an index exceeds its real axis (IndexError). */
extern "C" int mica_label_class_mask(const float* xyz, const uint8_t* is_class, int64_t n_atoms, float ox, float oy,
                                     float oz, int clip_x, int clip_y, int clip_z, int nz, int ny, int nx,
                                     int32_t* mask, int* status_oob, mica_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MICA_REQUIRE(mask && status_oob, "null pointer");
  MICA_REQUIRE(nz > 0 && ny > 0 && nx > 0, "empty grid");
  MICA_REQUIRE(n_atoms >= 0 && n_atoms < (1LL << 29), "atom count out of range");
  MICA_REQUIRE(clip_x >= 0 && clip_y >= 0 && clip_z >= 0, "negative clip bound");
  MICA_CUDA(cudaMemsetAsync(status_oob, 0, sizeof(int), st));
  MICA_CUDA(cudaMemsetAsync(mask, 0, sizeof(int32_t) * (size_t)nz * ny * nx, st));
  if (n_atoms == 0) return MICA_OK;
  MICA_REQUIRE(xyz && is_class, "null atom arrays");
  const unsigned grid = blocks_for(n_atoms, 256);
  class_mask_atoms_kernel<<<grid, 256, 0, st>>>(xyz, is_class, n_atoms, ox, oy, oz, clip_x, clip_y, clip_z, nz, ny, nx,
                                                mask, status_oob);
  MICA_LAUNCH_CHECK("class_mask_atoms_kernel");
  class_mask_finish_kernel<<<grid, 256, 0, st>>>(xyz, n_atoms, ox, oy, oz, clip_x, clip_y, clip_z, nz, ny, nx, mask);
  MICA_LAUNCH_CHECK("class_mask_finish_kernel");
  return MICA_OK;
}

extern "C" size_t mica_label_aa_mask_workspace_bytes(int nz, int ny, int nx) {
  return 3 * sizeof(int32_t) * (size_t)nz * ny * nx;
}

/* N3 amino-acid mask.  xyz: device float32 [R,3] C-alpha coordinates in file order; label: device int32 [R]
 * (1..20).  workspace: 3 int32 volumes. */
extern "C" int mica_label_aa_mask(const float* xyz, const int32_t* label, int64_t n_ca, float ox, float oy, float oz,
                                  int clip_x, int clip_y, int clip_z, int nz, int ny, int nx, void* workspace,
                                  size_t workspace_bytes, int32_t* mask, int* status_oob, mica_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MICA_REQUIRE(mask && status_oob && workspace, "null pointer");
  MICA_REQUIRE(nz > 0 && ny > 0 && nx > 0, "empty grid");
  MICA_REQUIRE(n_ca >= 0 && n_ca < (1LL << 31) - 1, "residue count out of range");
  if (workspace_bytes < mica_label_aa_mask_workspace_bytes(nz, ny, nx))
    return set_error(MICA_ERR_WORKSPACE, "amino-acid mask workspace too small");
  const size_t n_vox = (size_t)nz * ny * nx;
  int* a_all = (int*)workspace;
  int* last_on = a_all + n_vox;
  int* a_0 = last_on + n_vox;
  MICA_CUDA(cudaMemsetAsync(status_oob, 0, sizeof(int), st));
  MICA_CUDA(cudaMemsetAsync(mask, 0, sizeof(int32_t) * n_vox, st));
  if (n_ca == 0) return MICA_OK;
  MICA_REQUIRE(xyz && label, "null residue arrays");
  MICA_CUDA(cudaMemsetAsync(a_all, 0x7f, sizeof(int) * n_vox, st));
  MICA_CUDA(cudaMemsetAsync(last_on, 0, sizeof(int) * n_vox, st));
  MICA_CUDA(cudaMemsetAsync(a_0, 0x7f, sizeof(int) * n_vox, st));
  const unsigned grid = blocks_for(n_ca, 256);
  aa_mask_pass1_kernel<<<grid, 256, 0, st>>>(xyz, label, n_ca, ox, oy, oz, clip_x, clip_y, clip_z, nz, ny, nx, a_all,
                                             last_on, status_oob);
  MICA_LAUNCH_CHECK("aa_mask_pass1_kernel");
  aa_mask_pass2_kernel<<<grid, 256, 0, st>>>(xyz, label, n_ca, ox, oy, oz, clip_x, clip_y, clip_z, nz, ny, nx, last_on,
                                             a_0);
  MICA_LAUNCH_CHECK("aa_mask_pass2_kernel");
  aa_mask_pass3_kernel<<<grid, 256, 0, st>>>(xyz, n_ca, ox, oy, oz, clip_x, clip_y, clip_z, nz, ny, nx, a_all, last_on,
                                             a_0, mask);
  MICA_LAUNCH_CHECK("aa_mask_pass3_kernel");
  return MICA_OK;
}

/* N4 utils/dock_in_map.py:269 -- out = where(in < level, 0, in); in == out allowed */
extern "C" int mica_contour_threshold_f32(const float* in, float* out, int64_t n, float level, mica_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 0) return MICA_OK;
  MICA_REQUIRE(in && out, "null pointer");
  int64_t want = ceil_div64(n, 256 * 8);
  const int64_t cap = (int64_t)kNumSMs * 16;
  contour_threshold_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, st>>>(in, out, n, level);
  MICA_LAUNCH_CHECK("contour_threshold_kernel");
  return MICA_OK;
}

/* N4 utils/dock_in_map.py:330-352 -- zero `map` (device float32 [nz,ny,nx], in place) within `radius` of the
 * voxels of the atoms xyz (device float32 [A,3]).  voxel_xyz / origin_xyz: float32 header values. */
extern "C" int mica_zero_around_atoms(const float* xyz, int64_t n_atoms, const float origin_xyz[3],
                                      const float voxel_xyz[3], double radius, int nz, int ny, int nx, float* map,
                                      int* status_oob, mica_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MICA_REQUIRE(map && status_oob && origin_xyz && voxel_xyz, "null pointer");
  MICA_REQUIRE(nz > 0 && ny > 0 && nx > 0, "empty grid");
  MICA_REQUIRE(voxel_xyz[0] > 0 && voxel_xyz[1] > 0 && voxel_xyz[2] > 0, "voxel size must be positive");
  MICA_CUDA(cudaMemsetAsync(status_oob, 0, sizeof(int), st));
  if (n_atoms <= 0 || !(radius >= 0)) return MICA_OK;
  MICA_REQUIRE(xyz, "null atom array");
  // SciPy applies sampling = [vx, vy, vz] to axes (z, y, x)
  const double s0 = (double)voxel_xyz[0], s1 = (double)voxel_xyz[1], s2 = (double)voxel_xyz[2];
  const int rz = (int)(radius / s0) + 1, ry = (int)(radius / s1) + 1, rx = (int)(radius / s2) + 1;
  MICA_REQUIRE((int64_t)(2 * rz + 1) * (2 * ry + 1) * (2 * rx + 1) < (1LL << 30), "radius too large");
  zero_around_atoms_kernel<<<blocks_for(n_atoms * 32, 256), 256, 0, st>>>(
      xyz, n_atoms, origin_xyz[0], origin_xyz[1], origin_xyz[2], voxel_xyz[0], voxel_xyz[1], voxel_xyz[2], s0, s1, s2,
      radius, rz, ry, rx, nz, ny, nx, map, status_oob);
  MICA_LAUNCH_CHECK("zero_around_atoms_kernel");
  return MICA_OK;
}
