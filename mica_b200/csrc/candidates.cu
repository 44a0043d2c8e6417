// C-alpha candidates straight from the stitched volumes in HBM.
//
// Replaces the head of Solver.clustering (utils/modeler.py:762-860, reference root):
//   :767      np.where(CAProb > thr)                    -> ordered stream compaction
//   :768-770  Open3D DBSCAN on the voxel coordinates    -> lattice DBSCAN (bit volume + union-find)
//   :775-797  per-cluster backbone score                -> gather + per-label sums
//   :805-832  greedy NMS, best probability first        -> parallel greedy independent set
//   :837-860  3x3x3 weighted refinement + AA profile    -> one warp per pick
// so that the 20-channel amino_acid_probability volume (8.8 GB at 480^3) never has to cross
// PCIe: only the picks (a few thousand rows) do.
//
// Everything here is index work (bit-exact) except the cluster sums (float64 atomics vs the
// reference's float32 pairwise np.sum: ~1e-7 relative) and the refinement, which reproduces
// NumPy's operation order (pairwise 8-accumulator sum of the 27 weights, float64 centroid,
// float32 row-by-row profile) with explicit round-to-nearest intrinsics, i.e. no FMA contraction.
#include "common.cuh"

namespace mica {

int fill_zero(float* p, long long n, cudaStream_t st);  // af3.cu

// ------------------------------------------------------------------------------------------
// generic single-CTA exclusive scan of uint32 counts into int64 offsets (+ total)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
scan_u32_kernel(const uint32_t* __restrict__ in, long long m, long long* __restrict__ out_excl,
                long long* __restrict__ total) {
  __shared__ long long warp_excl[32];
  __shared__ long long chunk_total_s, carry_s;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (t == 0) carry_s = 0;
  __syncthreads();
  for (long long base = 0; base < m; base += 1024) {
    const long long i = base + t;
    const long long v = (i < m) ? (long long)in[i] : 0;
    long long inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const long long o = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += o;
    }
    if (lane == 31) warp_excl[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      const long long w = warp_excl[lane];
      long long winc = w;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const long long o = __shfl_up_sync(0xffffffffu, winc, d);
        if (lane >= d) winc += o;
      }
      warp_excl[lane] = winc - w;  // exclusive prefix of the warp totals
      if (lane == 31) chunk_total_s = winc;
    }
    __syncthreads();
    if (i < m) out_excl[i] = carry_s + warp_excl[warp] + inc - v;
    __syncthreads();  // everyone has read carry_s / warp_excl of this chunk
    if (t == 0) carry_s += chunk_total_s;
    __syncthreads();
  }
  if (t == 0 && total) *total = carry_s;
}

// ------------------------------------------------------------------------------------------
// :767 -- ordered compaction of the voxels above the threshold
// ------------------------------------------------------------------------------------------
constexpr int kTcThreads = 256;
constexpr int kTcIters = 4;
constexpr int kTcChunk = kTcThreads * 4 * kTcIters;  // 4096 voxels per CTA

__device__ __forceinline__ int load4_above(const float* __restrict__ v, long long i, long long n, float thr,
                                           bool hit[4]) {
  int c = 0;
  if (i + 3 < n) {
    const float4 q = ld_stream4(reinterpret_cast<const float4*>(v + i));
    hit[0] = q.x > thr; hit[1] = q.y > thr; hit[2] = q.z > thr; hit[3] = q.w > thr;
  } else {
#pragma unroll
    for (int e = 0; e < 4; ++e) hit[e] = (i + e < n) && (v[i + e] > thr);
  }
#pragma unroll
  for (int e = 0; e < 4; ++e) c += hit[e] ? 1 : 0;
  return c;
}

__global__ void __launch_bounds__(kTcThreads)
threshold_count_kernel(const float* __restrict__ v, long long n, float thr, uint32_t* __restrict__ block_counts) {
  __shared__ int warp_c[kTcThreads / 32];
  const long long base = (long long)blockIdx.x * kTcChunk;
  int c = 0;
  bool hit[4];
#pragma unroll
  for (int it = 0; it < kTcIters; ++it)
    c += load4_above(v, base + (long long)it * kTcThreads * 4 + 4 * threadIdx.x, n, thr, hit);
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0) warp_c[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
#pragma unroll
    for (int w = 0; w < kTcThreads / 32; ++w) s += warp_c[w];
    block_counts[blockIdx.x] = (uint32_t)s;
  }
}

__global__ void __launch_bounds__(kTcThreads)
threshold_write_kernel(const float* __restrict__ v, long long n, float thr, int Y, int Z,
                       const long long* __restrict__ block_offsets, long long cap,
                       long long* __restrict__ lin_out, int32_t* __restrict__ xyz_out) {
  __shared__ int warp_c[kTcThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long base = (long long)blockIdx.x * kTcChunk;
  long long run = block_offsets[blockIdx.x];
  for (int it = 0; it < kTcIters; ++it) {
    const long long i = base + (long long)it * kTcThreads * 4 + 4 * threadIdx.x;
    bool hit[4];
    const int c = load4_above(v, i, n, thr, hit);
    int inc = c;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int o = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += o;
    }
    if (lane == 31) warp_c[warp] = inc;
    __syncthreads();
    int before = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kTcThreads / 32; ++w) {
      const int wc = warp_c[w];
      if (w < warp) before += wc;
      total += wc;
    }
    long long pos = run + before + inc - c;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (hit[e]) {
        if (pos < cap) {
          const long long l = i + e;
          lin_out[pos] = l;
          if (xyz_out) {
            const long long xy = l / Z;
            xyz_out[3 * pos + 0] = (int32_t)(xy / Y);
            xyz_out[3 * pos + 1] = (int32_t)(xy % Y);
            xyz_out[3 * pos + 2] = (int32_t)(l - xy * Z);
          }
        }
        ++pos;
      }
    }
    run += total;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256)
gather_kernel(const float* __restrict__ vol, const long long* __restrict__ lin, long long n, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = vol[lin[i]];
}

// :775-786 -- per-label sum and count of the gathered backbone probabilities
__global__ void __launch_bounds__(256)
cluster_scores_kernel(const float* __restrict__ vals, const int32_t* __restrict__ labels, long long n, int n_labels,
                      double* __restrict__ sums, unsigned long long* __restrict__ counts) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int l = labels[i];
  if (l < 0 || l >= n_labels) return;
  atomicAdd(&sums[l], (double)vals[i]);
  atomicAdd(&counts[l], 1ULL);
}

// valid[i] = label_ok[labels[i]]  (:789-797)
__global__ void __launch_bounds__(256)
valid_points_kernel(const int32_t* __restrict__ labels, const uint8_t* __restrict__ label_ok, long long n,
                    int n_labels, uint8_t* __restrict__ valid) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int l = labels[i];
  valid[i] = (l >= 0 && l < n_labels) ? label_ok[l] : 0;
}

// :800-802 -- CAProb_clusted = zeros + CAProb at the valid points
__global__ void __launch_bounds__(256)
scatter_valid_kernel(const float* __restrict__ ca, const long long* __restrict__ lin, const uint8_t* __restrict__ valid,
                     long long n, float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !valid[i]) return;
  out[lin[i]] = ca[lin[i]];
}

// ------------------------------------------------------------------------------------------
// :805-832 -- greedy NMS as a parallel greedy independent set.
// The work volume w holds +p for an undecided valid point, -p once it is picked, 0 for
// suppressed points and everything else.  Priority: higher p first, ties by lower linear
// index (= the np.where order; the reference's unstable argsort leaves ties undefined).
// A point is picked once no undecided higher-priority point lies within the radius and no
// picked one does; it is suppressed as soon as a picked point lies within the radius.
// Transitions are final, so in-place updates only ever delay a decision: the fixed point is
// the sequential greedy result, whatever the schedule.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
nms_round_kernel(float* w, const long long* __restrict__ lin, const uint8_t* __restrict__ valid, long long n,
                 int X, int Y, int Z, int r, int radius2, int* __restrict__ remaining) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !valid[i]) return;
  const long long l = lin[i];
  volatile float* wv = w;
  const float p = wv[l];
  if (!(p > 0.f)) return;  // picked (< 0) or suppressed (0)
  const long long xy = l / Z;
  const int x = (int)(xy / Y), y = (int)(xy % Y), z = (int)(l - xy * Z);
  bool blocked = false;
  for (int dx = -r; dx <= r; ++dx) {
    const int xx = x + dx;
    if (xx < 0 || xx >= X) continue;
    for (int dy = -r; dy <= r; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= Y || dx * dx + dy * dy > radius2) continue;
      for (int dz = -r; dz <= r; ++dz) {
        const int zz = z + dz;
        if (zz < 0 || zz >= Z) continue;
        if (dx * dx + dy * dy + dz * dz > radius2) continue;
        if (dx == 0 && dy == 0 && dz == 0) continue;
        const long long nb = ((long long)xx * Y + yy) * Z + zz;
        const float q = wv[nb];
        if (q < 0.f) {  // a picked point within the radius: suppressed for good
          wv[l] = 0.f;
          return;
        }
        if (q > p || (q == p && nb < l)) blocked = true;
      }
    }
  }
  if (!blocked)
    wv[l] = -p;
  else
    *remaining = 1;
}

__global__ void __launch_bounds__(256)
nms_collect_kernel(const float* __restrict__ w, const long long* __restrict__ lin, const uint8_t* __restrict__ valid,
                   long long n, long long cap, long long* __restrict__ picked_lin, float* __restrict__ picked_p,
                   unsigned long long* __restrict__ count) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !valid[i]) return;
  const float p = w[lin[i]];
  if (p < 0.f) {
    const unsigned long long pos = atomicAdd(count, 1ULL);
    if ((long long)pos < cap) {
      picked_lin[pos] = lin[i];
      picked_p[pos] = -p;
    }
  }
}

// Order of the picks = the reference's argsort(-p): rank = number of picks with higher priority (higher p, ties by
// lower linear index).  Counting over all pairs would be m^2; instead the picks are bucketed by the top bits of p
// (descending), a scan gives each bucket its first rank, and the count runs inside the pick's own bucket only.
constexpr int kRankBuckets = 1 << 16;
__device__ __forceinline__ int rank_bucket(float p) {
  const uint32_t one = 0x3F800000u, bits = __float_as_uint(p);  // p > 0: the bit pattern is monotone in p
  if (bits >= one) return 0;
  const uint32_t b = (one - bits) >> 9;
  return b < (uint32_t)kRankBuckets ? (int)b : kRankBuckets - 1;
}

__global__ void __launch_bounds__(256)
nms_bucket_count_kernel(const float* __restrict__ picked_p, long long m, uint32_t* __restrict__ counts) {
  const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (a < m) atomicAdd(&counts[rank_bucket(picked_p[a])], 1u);
}

__global__ void __launch_bounds__(256)
nms_bucket_scatter_kernel(const long long* __restrict__ picked_lin, const float* __restrict__ picked_p, long long m,
                          const long long* __restrict__ bucket_first, uint32_t* __restrict__ fill,
                          long long* __restrict__ by_bucket_lin, float* __restrict__ by_bucket_p) {
  const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= m) return;
  const int b = rank_bucket(picked_p[a]);
  const long long pos = bucket_first[b] + atomicAdd(&fill[b], 1u);
  by_bucket_lin[pos] = picked_lin[a];
  by_bucket_p[pos] = picked_p[a];
}

__global__ void __launch_bounds__(256)
nms_rank_kernel(const long long* __restrict__ by_bucket_lin, const float* __restrict__ by_bucket_p, long long m,
                const long long* __restrict__ bucket_first, const uint32_t* __restrict__ counts, int Y, int Z,
                long long* __restrict__ sorted_lin, int32_t* __restrict__ sorted_xyz) {
  const long long a = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (a >= m) return;
  const float p = by_bucket_p[a];
  const long long l = by_bucket_lin[a];
  const int bkt = rank_bucket(p);
  const long long first = bucket_first[bkt], last = first + counts[bkt];
  long long rank = first;
  for (long long b = first; b < last; ++b) {
    const float q = by_bucket_p[b];
    rank += (q > p || (q == p && by_bucket_lin[b] < l)) ? 1 : 0;
  }
  sorted_lin[rank] = l;
  const long long xy = l / Z;
  sorted_xyz[3 * rank + 0] = (int32_t)(xy / Y);
  sorted_xyz[3 * rank + 1] = (int32_t)(xy % Y);
  sorted_xyz[3 * rank + 2] = (int32_t)(l - xy * Z);
}

// ------------------------------------------------------------------------------------------
// :837-860 -- one warp per pick: lanes 0..19 = amino-acid channels, lanes 20..22 = x,y,z
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
refine_kernel(const float* __restrict__ ca, const float* __restrict__ aa_prob, const float* __restrict__ aa_pred,
              int X, int Y, int Z, const long long* __restrict__ pick_lin, long long m,
              double* __restrict__ out_xyz, float* __restrict__ out_aaprob, float* __restrict__ out_aa,
              uint8_t* __restrict__ out_ok) {
  const long long pick = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (pick >= m) return;
  const long long l = pick_lin[pick];
  const long long xy = l / Z;
  const int c[3] = {(int)(xy / Y), (int)(xy % Y), (int)(l - xy * Z)};
  // a pick on the border makes the reference's slice short or empty -> IndexError -> skipped (:856)
  const bool inside = c[0] >= 1 && c[0] <= X - 2 && c[1] >= 1 && c[1] <= Y - 2 && c[2] >= 1 && c[2] <= Z - 2;
  if (lane == 0) out_ok[pick] = inside ? 1 : 0;
  if (!inside) return;
  float a[27];
#pragma unroll
  for (int k = 0; k < 27; ++k) {
    const int di = k / 9 - 1, dj = (k / 3) % 3 - 1, dk = k % 3 - 1;
    a[k] = ca[((long long)(c[0] + di) * Y + (c[1] + dj)) * Z + (c[2] + dk)];
  }
  // np.sum over the 27 ravelled float32 values: NumPy's pairwise kernel (8 running sums over the first
  // 24, combined as a tree, then the last three one by one)
  float r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = a[j];
#pragma unroll
  for (int i = 8; i < 24; i += 8)
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a[i + j]);
  float s = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                      __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
  s = __fadd_rn(s, a[24]);
  s = __fadd_rn(s, a[25]);
  s = __fadd_rn(s, a[26]);
  if (lane < 20) {
    const long long chan = (long long)X * Y * Z;
    const float* src = aa_prob + lane * chan;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 27; ++k) {
      const int di = k / 9 - 1, dj = (k / 3) % 3 - 1, dk = k % 3 - 1;
      const float wk = __fdiv_rn(a[k], s);
      const float term = __fmul_rn(src[((long long)(c[0] + di) * Y + (c[1] + dj)) * Z + (c[2] + dk)], wk);
      acc = (k == 0) ? term : __fadd_rn(acc, term);
    }
    out_aaprob[pick * 20 + lane] = acc;
  } else if (lane < 23) {
    const int ax = lane - 20;
    double acc = 0.0;
#pragma unroll
    for (int k = 0; k < 27; ++k) {
      const int d[3] = {k / 9 - 1, (k / 3) % 3 - 1, k % 3 - 1};
      const float wk = __fdiv_rn(a[k], s);
      acc = __dadd_rn(acc, __dmul_rn((double)(c[ax] + d[ax]), (double)wk));
    }
    out_xyz[pick * 3 + ax] = acc;
  }
  // CA_cands_AA = AAPred[round(coord)] (:858-860); every lane recomputes the three coordinates
  if (lane == 23) {
    long long idx = 0;
    for (int ax = 0; ax < 3; ++ax) {
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < 27; ++k) {
        const int d[3] = {k / 9 - 1, (k / 3) % 3 - 1, k % 3 - 1};
        const float wk = __fdiv_rn(a[k], s);
        acc = __dadd_rn(acc, __dmul_rn((double)(c[ax] + d[ax]), (double)wk));
      }
      long long q = (long long)rint(acc);  // np.round: half to even
      const int dim = ax == 0 ? X : (ax == 1 ? Y : Z);
      q = q < 0 ? 0 : (q > dim - 1 ? dim - 1 : q);
      idx = idx * dim + q;
    }
    out_aa[pick] = aa_pred[idx];
  }
}

// ------------------------------------------------------------------------------------------
// :768-770 -- DBSCAN on lattice points.  The points are distinct voxels, so the eps-ball is a
// fixed set of lattice offsets: occupancy lives in a bit volume (bits along Z), a point is a
// core point when popcount(ball) >= min_points (itself included, closed ball: d^2 <= eps^2),
// core points within eps of each other are merged by a lock-free union-find whose roots are the
// lowest point index of the component (= the order in which sequential DBSCAN opens clusters),
// and a border point takes the lowest-numbered cluster that has a core point within eps
// (= the first cluster to reach it).  Noise = -1.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ int isqrt_floor(int v) {
  int h = (int)sqrtf((float)v);
  while (h * h > v) --h;
  while ((h + 1) * (h + 1) <= v) ++h;
  return h;
}

struct Lattice {
  int X, Y, Z, Wz;  // Wz = 32-bit words per (x,y) row
};

__device__ __forceinline__ void lin_to_xyz(long long l, const Lattice& g, int& x, int& y, int& z) {
  const long long xy = l / g.Z;
  x = (int)(xy / g.Y);
  y = (int)(xy % g.Y);
  z = (int)(l - xy * g.Z);
}

// calls f(xx, yy, word_index_in_row, bits) for every non-empty 32-bit word of `bitvol` inside the ball;
// f returns true to stop early.  The (dx,dy) columns of the ball are visited first, first+step, ... so that
// the lanes of a warp can share one point (first = lane, step = 32) or a thread can own it (0, 1).
template <class F>
__device__ __forceinline__ void for_ball_words(const uint32_t* __restrict__ bitvol, const Lattice& g, int x, int y,
                                               int z, int R, int eps2, int first, int step, F f) {
  const int side = 2 * R + 1;
  for (int t = first; t < side * side; t += step) {
    const int dx = t / side - R, dy = t % side - R;
    const int xx = x + dx, yy = y + dy;
    const int rem = eps2 - dx * dx - dy * dy;
    if (xx < 0 || xx >= g.X || yy < 0 || yy >= g.Y || rem < 0) continue;
    const int hz = isqrt_floor(rem);
    const int z0 = max(0, z - hz), z1 = min(g.Z - 1, z + hz);
    const uint32_t* row = bitvol + ((long long)xx * g.Y + yy) * g.Wz;
    const int w0 = z0 >> 5, w1 = z1 >> 5;
    for (int w = w0; w <= w1; ++w) {
      uint32_t mask = 0xffffffffu;
      if (w == w0) mask &= 0xffffffffu << (z0 & 31);
      if (w == w1) mask &= 0xffffffffu >> (31 - (z1 & 31));
      const uint32_t bits = row[w] & mask;
      if (bits && f(xx, yy, w, bits)) return;
    }
  }
}

__global__ void __launch_bounds__(256)
dbscan_mark_kernel(const long long* __restrict__ lin, long long n, Lattice g, uint32_t* __restrict__ occ,
                   int32_t* __restrict__ slot, int32_t* __restrict__ parent) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int x, y, z;
  lin_to_xyz(lin[i], g, x, y, z);
  atomicOr(&occ[((long long)x * g.Y + y) * g.Wz + (z >> 5)], 1u << (z & 31));
  slot[lin[i]] = (int32_t)i;
  parent[i] = (int32_t)i;
}

__global__ void __launch_bounds__(256)
dbscan_core_kernel(const long long* __restrict__ lin, long long n, Lattice g, const uint32_t* __restrict__ occ,
                   int R, int eps2, int min_points, uint8_t* __restrict__ is_core, uint32_t* __restrict__ coreocc) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int x, y, z;
  lin_to_xyz(lin[i], g, x, y, z);
  int count = 0;
  for_ball_words(occ, g, x, y, z, R, eps2, 0, 1, [&](int, int, int, uint32_t bits) {
    count += __popc(bits);
    return count >= min_points;
  });
  const bool core = count >= min_points;
  is_core[i] = core ? 1 : 0;
  if (core) atomicOr(&coreocc[((long long)x * g.Y + y) * g.Wz + (z >> 5)], 1u << (z & 31));
}

__device__ __forceinline__ int uf_find(const int32_t* parent, int i) {
  const volatile int32_t* p = parent;
  int q = p[i];
  while (q != i) {
    i = q;
    q = p[i];
  }
  return i;
}

// find with path halving: every visited node is re-pointed to its grandparent.  Racing writers only ever
// store an ancestor (a hooked root never becomes a root again and hooks go to smaller indices), so the
// forest stays valid; roots themselves are written by atomicCAS only.
__device__ __forceinline__ int uf_find_compress(int32_t* parent, int i) {
  volatile int32_t* p = parent;
  int q = p[i];
  while (q != i) {
    const int g = p[q];
    if (g != q) p[i] = g;
    i = q;
    q = g;
  }
  return i;
}

// merges the trees of roots a and b; returns the root of the merged tree (the smaller index)
__device__ __forceinline__ int uf_union_roots(int32_t* parent, int a, int b) {
  while (a != b) {
    if (a < b) {
      const int t = a;
      a = b;
      b = t;
    }
    // a > b: hang root a under the smaller root b; roots only ever point to smaller indices
    const int old = atomicCAS(&parent[a], a, b);
    if (old == a) return b;
    a = uf_find_compress(parent, old);  // somebody else hooked a meanwhile: continue from its new root
    b = uf_find_compress(parent, b);
  }
  return a;
}

// one warp per core point: the lanes take the (dx,dy) columns of its ball in turn
__global__ void __launch_bounds__(256)
dbscan_union_kernel(const long long* __restrict__ lin, long long n, Lattice g, const uint32_t* __restrict__ coreocc,
                    const int32_t* __restrict__ slot, const uint8_t* __restrict__ is_core, int R, int eps2,
                    int32_t* parent) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n || !is_core[i]) return;
  int x, y, z;
  lin_to_xyz(lin[i], g, x, y, z);
  int my_root = uf_find_compress(parent, (int)i);
  for_ball_words(coreocc, g, x, y, z, R, eps2, lane, 32, [&](int xx, int yy, int w, uint32_t bits) {
    while (bits) {
      const int b = __ffs(bits) - 1;
      bits &= bits - 1;
      const int j = slot[((long long)xx * g.Y + yy) * g.Z + (w * 32 + b)];
      if (j < 0 || j >= (int)i) continue;  // each pair once, from its higher index
      const int rj = uf_find_compress(parent, j);
      if (rj != my_root) my_root = uf_union_roots(parent, uf_find_compress(parent, my_root), rj);
    }
    return false;
  });
}

// root of every core point; lowest root among the core points within eps for the others (-1: noise)
__global__ void __launch_bounds__(256)
dbscan_root_kernel(const long long* __restrict__ lin, long long n, Lattice g, const uint32_t* __restrict__ coreocc,
                   const int32_t* __restrict__ slot, const uint8_t* __restrict__ is_core, int R, int eps2,
                   const int32_t* __restrict__ parent, int32_t* __restrict__ root_of, uint32_t* __restrict__ is_root) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (is_core[i]) {
    const int r = uf_find(parent, (int)i);
    root_of[i] = r;
    if (r == (int)i) is_root[i] = 1u;
    return;
  }
  int x, y, z;
  lin_to_xyz(lin[i], g, x, y, z);
  int best = 0x7fffffff;
  for_ball_words(coreocc, g, x, y, z, R, eps2, 0, 1, [&](int xx, int yy, int w, uint32_t bits) {
    while (bits) {
      const int b = __ffs(bits) - 1;
      bits &= bits - 1;
      const int j = slot[((long long)xx * g.Y + yy) * g.Z + (w * 32 + b)];
      if (j >= 0) best = min(best, uf_find(parent, j));
    }
    return false;
  });
  root_of[i] = best == 0x7fffffff ? -1 : best;
}

__global__ void __launch_bounds__(256)
dbscan_label_kernel(const int32_t* __restrict__ root_of, const long long* __restrict__ cluster_of_root, long long n,
                    int32_t* __restrict__ labels) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = root_of[i];
  labels[i] = r < 0 ? -1 : (int32_t)cluster_of_root[r];
}

static inline unsigned grid_for(long long n, int threads) { return (unsigned)ceil_div64(n > 0 ? n : 1, threads); }

static inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace mica

using namespace mica;

// ==========================================================================================
// C ABI
// ==========================================================================================
extern "C" size_t mica_cand_threshold_workspace_bytes(int64_t n_vox) {
  const int64_t nb = ceil_div64(n_vox > 0 ? n_vox : 1, kTcChunk);
  return align256((size_t)nb * sizeof(uint32_t)) + align256((size_t)nb * sizeof(long long));
}

extern "C" int mica_cand_threshold_count(const float* vol, int64_t n_vox, float thr, void* workspace,
                                         size_t workspace_bytes, int64_t* count_dev, mica_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MICA_REQUIRE(vol && workspace && count_dev, "null pointer");
  MICA_REQUIRE(n_vox > 0, "empty volume");
  MICA_REQUIRE(((uintptr_t)vol & 15) == 0, "volume must be 16-byte aligned");
  if (workspace_bytes < mica_cand_threshold_workspace_bytes(n_vox))
    return set_error(MICA_ERR_WORKSPACE, "threshold workspace too small");
  const int64_t nb = ceil_div64(n_vox, kTcChunk);
  uint32_t* counts = (uint32_t*)workspace;
  long long* offsets = (long long*)((char*)workspace + align256((size_t)nb * sizeof(uint32_t)));
  threshold_count_kernel<<<(unsigned)nb, kTcThreads, 0, st>>>(vol, n_vox, thr, counts);
  MICA_LAUNCH_CHECK("threshold_count_kernel");
  scan_u32_kernel<<<1, 1024, 0, st>>>(counts, nb, offsets, (long long*)count_dev);
  MICA_LAUNCH_CHECK("scan_u32_kernel");
  return MICA_OK;
}

extern "C" int mica_cand_threshold_write(const float* vol, int X, int Y, int Z, float thr, const void* workspace,
                                         int64_t* lin_out, int32_t* xyz_out, int64_t cap, mica_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MICA_REQUIRE(vol && workspace && lin_out, "null pointer");
  MICA_REQUIRE(X > 0 && Y > 0 && Z > 0, "empty volume");
  const int64_t n_vox = (int64_t)X * Y * Z;
  const int64_t nb = ceil_div64(n_vox, kTcChunk);
  const long long* offsets = (const long long*)((const char*)workspace + align256((size_t)nb * sizeof(uint32_t)));
  threshold_write_kernel<<<(unsigned)nb, kTcThreads, 0, st>>>(vol, n_vox, thr, Y, Z, offsets, cap,
                                                              (long long*)lin_out, xyz_out);
  MICA_LAUNCH_CHECK("threshold_write_kernel");
  return MICA_OK;
}

extern "C" int mica_gather_f32(const float* vol, const int64_t* lin, int64_t n, float* out, mica_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 0) return MICA_OK;
  MICA_REQUIRE(vol && lin && out, "null pointer");
  gather_kernel<<<grid_for(n, 256), 256, 0, st>>>(vol, (const long long*)lin, n, out);
  MICA_LAUNCH_CHECK("gather_kernel");
  return MICA_OK;
}

extern "C" int mica_cand_cluster_scores(const float* vals, const int32_t* labels, int64_t n, int n_labels,
                                        double* sums, int64_t* counts, mica_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MICA_REQUIRE(n_labels >= 0, "negative label count");
  if (n_labels == 0) return MICA_OK;
  MICA_REQUIRE(sums && counts, "null pointer");
  MICA_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * n_labels, st));
  MICA_CUDA(cudaMemsetAsync(counts, 0, sizeof(int64_t) * n_labels, st));
  if (n <= 0) return MICA_OK;
  MICA_REQUIRE(vals && labels, "null pointer");
  cluster_scores_kernel<<<grid_for(n, 256), 256, 0, st>>>(vals, labels, n, n_labels, sums,
                                                          (unsigned long long*)counts);
  MICA_LAUNCH_CHECK("cluster_scores_kernel");
  return MICA_OK;
}

extern "C" int mica_cand_valid_points(const int32_t* labels, const uint8_t* label_ok, int64_t n, int n_labels,
                                      uint8_t* valid, mica_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 0) return MICA_OK;
  MICA_REQUIRE(labels && valid && (n_labels == 0 || label_ok), "null pointer");
  valid_points_kernel<<<grid_for(n, 256), 256, 0, st>>>(labels, label_ok, n, n_labels, valid);
  MICA_LAUNCH_CHECK("valid_points_kernel");
  return MICA_OK;
}

extern "C" int mica_cand_clustered_volume(const float* ca, int64_t n_vox, const int64_t* lin, const uint8_t* valid,
                                          int64_t n, float* out, mica_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MICA_REQUIRE(ca && out && n_vox > 0, "null pointer / empty volume");
  int rc = fill_zero(out, n_vox, st);
  if (rc) return rc;
  if (n <= 0) return MICA_OK;
  MICA_REQUIRE(lin && valid, "null pointer");
  scatter_valid_kernel<<<grid_for(n, 256), 256, 0, st>>>(ca, (const long long*)lin, valid, n, out);
  MICA_LAUNCH_CHECK("scatter_valid_kernel");
  return MICA_OK;
}

/* work: device float32 [X*Y*Z]; on return it holds -p at the picks (and 0 elsewhere).  Runs rounds until
 * no point is undecided (synchronises the stream every 8 rounds to read one flag). */
extern "C" int mica_cand_nms(const float* ca, int X, int Y, int Z, const int64_t* lin, const uint8_t* valid,
                             int64_t n, int nms_radius_sq, float* work, int32_t* flag_dev, int* rounds_out,
                             mica_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MICA_REQUIRE(ca && work && flag_dev, "null pointer");
  MICA_REQUIRE(X > 0 && Y > 0 && Z > 0 && nms_radius_sq >= 0, "bad arguments");
  const int64_t n_vox = (int64_t)X * Y * Z;
  int rc = fill_zero(work, n_vox, st);
  if (rc) return rc;
  int rounds = 0;
  if (n > 0) {
    MICA_REQUIRE(lin && valid, "null pointer");
    scatter_valid_kernel<<<grid_for(n, 256), 256, 0, st>>>(ca, (const long long*)lin, valid, n, work);
    MICA_LAUNCH_CHECK("scatter_valid_kernel");
    int r = 0;
    while ((r + 1) * (r + 1) <= nms_radius_sq) ++r;
    const int kBatch = 8, kMaxRounds = 1 << 16;
    int remaining = 1;
    while (remaining && rounds < kMaxRounds) {
      for (int k = 0; k < kBatch; ++k) {
        if (k == kBatch - 1) MICA_CUDA(cudaMemsetAsync(flag_dev, 0, sizeof(int32_t), st));
        nms_round_kernel<<<grid_for(n, 256), 256, 0, st>>>(work, (const long long*)lin, valid, n, X, Y, Z, r,
                                                           nms_radius_sq, flag_dev);
        MICA_LAUNCH_CHECK("nms_round_kernel");
        ++rounds;
      }
      MICA_CUDA(cudaMemcpyAsync(&remaining, flag_dev, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
      MICA_CUDA(cudaStreamSynchronize(st));
    }
    if (remaining) return set_error(MICA_ERR_CUDA, "NMS did not converge in %d rounds", rounds);
  }
  if (rounds_out) *rounds_out = rounds;
  return MICA_OK;
}

extern "C" size_t mica_cand_picks_workspace_bytes(int64_t cap) {
  const size_t c = (size_t)(cap > 0 ? cap : 1);
  return align256(kRankBuckets * sizeof(uint32_t)) * 2 + align256(kRankBuckets * sizeof(long long)) +
         2 * (align256(c * sizeof(long long)) + align256(c * sizeof(float)));
}

/* picks in the reference's order (best probability first).  workspace: mica_cand_picks_workspace_bytes(cap). */
extern "C" int mica_cand_nms_picks(const float* work, int Y, int Z, const int64_t* lin, const uint8_t* valid,
                                   int64_t n, int64_t cap, void* workspace, size_t workspace_bytes,
                                   int64_t* n_picks_dev, int64_t* sorted_lin, int32_t* sorted_xyz,
                                   mica_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MICA_REQUIRE(work && n_picks_dev, "null pointer");
  MICA_CUDA(cudaMemsetAsync(n_picks_dev, 0, sizeof(int64_t), st));
  if (n <= 0 || cap <= 0) return MICA_OK;
  MICA_REQUIRE(lin && valid && workspace && sorted_lin && sorted_xyz, "null pointer");
  if (workspace_bytes < mica_cand_picks_workspace_bytes(cap))
    return set_error(MICA_ERR_WORKSPACE, "picks workspace too small");
  char* ws = (char*)workspace;
  uint32_t* counts = (uint32_t*)ws;
  uint32_t* fill = (uint32_t*)(ws + align256(kRankBuckets * sizeof(uint32_t)));
  long long* first = (long long*)(ws + 2 * align256(kRankBuckets * sizeof(uint32_t)));
  char* q = (char*)first + align256(kRankBuckets * sizeof(long long));
  long long* picked_lin = (long long*)q; q += align256((size_t)cap * sizeof(long long));
  float* picked_p = (float*)q; q += align256((size_t)cap * sizeof(float));
  long long* bucket_lin = (long long*)q; q += align256((size_t)cap * sizeof(long long));
  float* bucket_p = (float*)q;
  nms_collect_kernel<<<grid_for(n, 256), 256, 0, st>>>(work, (const long long*)lin, valid, n, cap, picked_lin,
                                                       picked_p, (unsigned long long*)n_picks_dev);
  MICA_LAUNCH_CHECK("nms_collect_kernel");
  int64_t m = 0;
  MICA_CUDA(cudaMemcpyAsync(&m, n_picks_dev, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  MICA_CUDA(cudaStreamSynchronize(st));
  if (m > cap) return set_error(MICA_ERR_WORKSPACE, "%lld picks exceed the capacity %lld", (long long)m, (long long)cap);
  if (m > 0) {
    MICA_CUDA(cudaMemsetAsync(counts, 0, 2 * align256(kRankBuckets * sizeof(uint32_t)), st));  // counts + fill
    nms_bucket_count_kernel<<<grid_for(m, 256), 256, 0, st>>>(picked_p, m, counts);
    MICA_LAUNCH_CHECK("nms_bucket_count_kernel");
    scan_u32_kernel<<<1, 1024, 0, st>>>(counts, kRankBuckets, first, nullptr);
    MICA_LAUNCH_CHECK("scan_u32_kernel");
    nms_bucket_scatter_kernel<<<grid_for(m, 256), 256, 0, st>>>(picked_lin, picked_p, m, first, fill, bucket_lin,
                                                                bucket_p);
    MICA_LAUNCH_CHECK("nms_bucket_scatter_kernel");
    nms_rank_kernel<<<grid_for(m, 256), 256, 0, st>>>(bucket_lin, bucket_p, m, first, counts, Y, Z,
                                                      (long long*)sorted_lin, sorted_xyz);
    MICA_LAUNCH_CHECK("nms_rank_kernel");
  }
  return MICA_OK;
}

extern "C" int mica_cand_refine(const float* ca, const float* aa_prob, const float* aa_pred, int X, int Y, int Z,
                                const int64_t* pick_lin, int64_t m, double* out_xyz, float* out_aaprob,
                                float* out_aa, uint8_t* out_ok, mica_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (m <= 0) return MICA_OK;
  MICA_REQUIRE(ca && aa_prob && aa_pred && pick_lin && out_xyz && out_aaprob && out_aa && out_ok, "null pointer");
  MICA_REQUIRE(X > 0 && Y > 0 && Z > 0, "empty volume");
  refine_kernel<<<grid_for(m * 32, 128), 128, 0, st>>>(ca, aa_prob, aa_pred, X, Y, Z, (const long long*)pick_lin, m,
                                                       out_xyz, out_aaprob, out_aa, out_ok);
  MICA_LAUNCH_CHECK("refine_kernel");
  return MICA_OK;
}

// ---- DBSCAN ------------------------------------------------------------------------------
namespace {
struct DbscanLayout {
  size_t occ, coreocc, slot, parent, root_of, is_root, cluster_of_root, is_core, total;
};
DbscanLayout dbscan_layout(int X, int Y, int Z, int64_t n) {
  DbscanLayout L;
  const size_t words = (size_t)X * Y * ((Z + 31) / 32);
  const size_t n_vox = (size_t)X * Y * Z;
  const size_t nn = (size_t)(n > 0 ? n : 1);
  size_t off = 0;
  L.occ = off; off += align256(words * 4);
  L.coreocc = off; off += align256(words * 4);
  L.slot = off; off += align256(n_vox * 4);
  L.parent = off; off += align256(nn * 4);
  L.root_of = off; off += align256(nn * 4);
  L.is_root = off; off += align256(nn * 4);
  L.cluster_of_root = off; off += align256(nn * 8);
  L.is_core = off; off += align256(nn);
  L.total = off;
  return L;
}
}  // namespace

extern "C" size_t mica_dbscan_workspace_bytes(int X, int Y, int Z, int64_t n_points) {
  return dbscan_layout(X, Y, Z, n_points).total;
}

/* lin: device int64 [n], distinct voxels of an (X,Y,Z) C-order volume in ascending order (the output of
 * mica_cand_threshold_write).  eps_sq = floor(eps^2).  labels: device int32 [n]; n_clusters_dev: device int64. */
extern "C" int mica_dbscan_lattice(const int64_t* lin, int64_t n, int X, int Y, int Z, int eps_sq, int min_points,
                                   void* workspace, size_t workspace_bytes, int32_t* labels,
                                   int64_t* n_clusters_dev, mica_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MICA_REQUIRE(n_clusters_dev, "null pointer");
  MICA_CUDA(cudaMemsetAsync(n_clusters_dev, 0, sizeof(int64_t), st));
  if (n <= 0) return MICA_OK;
  MICA_REQUIRE(lin && labels && workspace, "null pointer");
  MICA_REQUIRE(X > 0 && Y > 0 && Z > 0 && eps_sq >= 0 && min_points >= 1, "bad arguments");
  MICA_REQUIRE(n < (1LL << 31), "too many points");
  const DbscanLayout L = dbscan_layout(X, Y, Z, n);
  if (workspace_bytes < L.total) return set_error(MICA_ERR_WORKSPACE, "dbscan workspace too small");
  char* ws = (char*)workspace;
  uint32_t* occ = (uint32_t*)(ws + L.occ);
  uint32_t* coreocc = (uint32_t*)(ws + L.coreocc);
  int32_t* slot = (int32_t*)(ws + L.slot);
  int32_t* parent = (int32_t*)(ws + L.parent);
  int32_t* root_of = (int32_t*)(ws + L.root_of);
  uint32_t* is_root = (uint32_t*)(ws + L.is_root);
  long long* cluster_of_root = (long long*)(ws + L.cluster_of_root);
  uint8_t* is_core = (uint8_t*)(ws + L.is_core);
  Lattice g{X, Y, Z, (Z + 31) / 32};
  int R = 0;
  while ((R + 1) * (R + 1) <= eps_sq) ++R;
  MICA_CUDA(cudaMemsetAsync(occ, 0, L.slot - L.occ, st));             // occ + coreocc
  MICA_CUDA(cudaMemsetAsync(slot, 0xff, (size_t)X * Y * Z * 4, st));  // -1
  MICA_CUDA(cudaMemsetAsync(is_root, 0, (size_t)n * 4, st));
  const unsigned grid = grid_for(n, 256);
  const long long* l = (const long long*)lin;
  dbscan_mark_kernel<<<grid, 256, 0, st>>>(l, n, g, occ, slot, parent);
  MICA_LAUNCH_CHECK("dbscan_mark_kernel");
  dbscan_core_kernel<<<grid, 256, 0, st>>>(l, n, g, occ, R, eps_sq, min_points, is_core, coreocc);
  MICA_LAUNCH_CHECK("dbscan_core_kernel");
  dbscan_union_kernel<<<grid_for(n * 32, 256), 256, 0, st>>>(l, n, g, coreocc, slot, is_core, R, eps_sq, parent);
  MICA_LAUNCH_CHECK("dbscan_union_kernel");
  dbscan_root_kernel<<<grid, 256, 0, st>>>(l, n, g, coreocc, slot, is_core, R, eps_sq, parent, root_of, is_root);
  MICA_LAUNCH_CHECK("dbscan_root_kernel");
  scan_u32_kernel<<<1, 1024, 0, st>>>(is_root, n, cluster_of_root, (long long*)n_clusters_dev);
  MICA_LAUNCH_CHECK("scan_u32_kernel");
  dbscan_label_kernel<<<grid, 256, 0, st>>>(root_of, cluster_of_root, n, labels);
  MICA_LAUNCH_CHECK("dbscan_label_kernel");
  return MICA_OK;
}

// ==========================================================================================
// utils/modeler.py:862-899 -- the neighbour graph of the picks (rest of Solver.clustering)
// ==========================================================================================
namespace mica {

// cand_self_dis[a,b] = np.linalg.norm(c[a] - c[b]) (:862, calc_dis :174-181): sqrt((dx^2 + dy^2) + dz^2) in
// float64; neigh_mat[a,b] (:875-886) for 2 <= dis <= 6, else 0.
__global__ void __launch_bounds__(256)
neighbor_graph_kernel(const double* __restrict__ xyz, long long m, const float* __restrict__ bb, int X, int Y, int Z,
                      double* __restrict__ dis_out, double* __restrict__ neigh_out) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long a = blockIdx.y;
  if (b >= m) return;
  const double ax = xyz[3 * a], ay = xyz[3 * a + 1], az = xyz[3 * a + 2];
  const double bx = xyz[3 * b], by = xyz[3 * b + 1], bz = xyz[3 * b + 2];
  const double dx = __dsub_rn(ax, bx), dy = __dsub_rn(ay, by), dz = __dsub_rn(az, bz);
  const double d = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)));
  dis_out[a * m + b] = d;
  double score = 0.0;
  if (d <= 6.0 && d >= 2.0) {
    // BB_dens: float32 sum of the backbone probability at the four interior fifths of the segment a -> b
    float dens = 0.f;
#pragma unroll
    for (int j = 1; j <= 4; ++j) {
      const double wb = (double)j / 5.0, wa = (double)(5 - j) / 5.0;   // python: j/5, (5-j)/5
      const long long cx = (long long)rint(__dadd_rn(__dmul_rn(wb, bx), __dmul_rn(wa, ax)));
      const long long cy = (long long)rint(__dadd_rn(__dmul_rn(wb, by), __dmul_rn(wa, ay)));
      const long long cz = (long long)rint(__dadd_rn(__dmul_rn(wb, bz), __dmul_rn(wa, az)));
      const long long ix = cx < 0 ? cx + X : cx, iy = cy < 0 ? cy + Y : cy, iz = cz < 0 ? cz + Z : cz;  // numpy wrap
      float v = 0.f;
      if (ix >= 0 && ix < X && iy >= 0 && iy < Y && iz >= 0 && iz < Z) v = bb[(ix * Y + iy) * Z + iz];
      dens = __fadd_rn(dens, v);
    }
    const float quarter = __fdiv_rn(dens, 4.0f);                      // BB_dens / 4 stays float32
    const double t = __dsub_rn(fabs(__dsub_rn(d, 3.8)), 0.5);         // abs(dis - 3.8) - 0.5
    if (t <= 0.0) {
      // max(0, t) is the python int 0 -> dis_score is the python float 1.0 -> the sum is formed in FLOAT32
      score = (double)__fdiv_rn(__fadd_rn(1.0f, quarter), 2.0f);
    } else {
      // np.float64 path: dis_score = max(0, 1 - t/2) (> 0 for every d in [2, 6])
      const double ds = __dsub_rn(1.0, __ddiv_rn(t, 2.0));
      score = ds > 0.0 ? __ddiv_rn(__dadd_rn(ds, (double)quarter), 2.0) : (double)__fdiv_rn(quarter, 2.0f);
    }
  }
  neigh_out[a * m + b] = score;
}

// :889-897 -- the two best-scoring neighbours of every pick: best[a] = (first, second), -1 where the score is 0.
// Ascending stable order semantics: among equal scores the higher index counts as larger.
__global__ void __launch_bounds__(256)
best_neighbors_kernel(const double* __restrict__ neigh, long long m, int32_t* __restrict__ best) {
  const long long a = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (a >= m) return;
  double v1 = -1.0, v2 = -1.0;  // scores are >= 0
  int i1 = -1, i2 = -1;
  const double* row = neigh + a * m;
  for (long long b = lane; b < m; b += 32) {
    const double v = row[b];
    if (v > v1 || (v == v1 && (int)b > i1)) {
      v2 = v1; i2 = i1; v1 = v; i1 = (int)b;
    } else if (v > v2 || (v == v2 && (int)b > i2)) {
      v2 = v; i2 = (int)b;
    }
  }
  // merge the per-lane top-2 lists
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) {
    const double o1 = __shfl_xor_sync(0xffffffffu, v1, d), o2 = __shfl_xor_sync(0xffffffffu, v2, d);
    const int j1 = __shfl_xor_sync(0xffffffffu, i1, d), j2 = __shfl_xor_sync(0xffffffffu, i2, d);
    double c[4] = {v1, v2, o1, o2};
    int ci[4] = {i1, i2, j1, j2};
    double n1 = -1.0, n2 = -1.0;
    int k1 = -1, k2 = -1;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      if (ci[t] < 0 || ci[t] == k1 || ci[t] == k2) continue;
      if (c[t] > n1 || (c[t] == n1 && ci[t] > k1)) {
        n2 = n1; k2 = k1; n1 = c[t]; k1 = ci[t];
      } else if (c[t] > n2 || (c[t] == n2 && ci[t] > k2)) {
        n2 = c[t]; k2 = ci[t];
      }
    }
    v1 = n1; v2 = n2; i1 = k1; i2 = k2;
  }
  if (lane == 0) {
    best[2 * a + 0] = (i1 >= 0 && v1 != 0.0) ? i1 : -1;
    best[2 * a + 1] = (i2 >= 0 && v2 != 0.0) ? i2 : -1;
  }
}

// :866-873 -- per pick the ascending list of picks within max_dis (one warp per row, ballot-ordered append)
__global__ void __launch_bounds__(256)
neighbor_lists_kernel(const double* __restrict__ dis, long long m, double max_dis, int cap, int32_t* __restrict__ idx,
                      int32_t* __restrict__ count) {
  const long long a = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (a >= m) return;
  const double* row = dis + a * m;
  int n = 0;
  for (long long b0 = 0; b0 < m; b0 += 32) {
    const long long b = b0 + lane;
    const bool hit = b < m && row[b] <= max_dis;
    const unsigned mask = __ballot_sync(0xffffffffu, hit);
    if (hit) {
      const int pos = n + __popc(mask & ((1u << lane) - 1u));
      if (pos < cap) idx[a * cap + pos] = (int32_t)b;
    }
    n += __popc(mask);
  }
  if (lane == 0) count[a] = n;
}

}  // namespace mica

extern "C" int mica_cand_neighbor_graph(const double* xyz, int64_t m, const float* bb, int X, int Y, int Z,
                                        double* dis_out, double* neigh_out, mica_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (m <= 0) return MICA_OK;
  MICA_REQUIRE(xyz && bb && dis_out && neigh_out, "null pointer");
  MICA_REQUIRE(X > 0 && Y > 0 && Z > 0 && m <= 65535, "bad arguments (at most 65535 picks)");
  dim3 grid(grid_for(m, 256), (unsigned)m);
  neighbor_graph_kernel<<<grid, 256, 0, st>>>(xyz, m, bb, X, Y, Z, dis_out, neigh_out);
  MICA_LAUNCH_CHECK("neighbor_graph_kernel");
  return MICA_OK;
}

extern "C" int mica_cand_best_neighbors(const double* neigh, int64_t m, int32_t* best, mica_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (m <= 0) return MICA_OK;
  MICA_REQUIRE(neigh && best, "null pointer");
  best_neighbors_kernel<<<grid_for(m * 32, 256), 256, 0, st>>>(neigh, m, best);
  MICA_LAUNCH_CHECK("best_neighbors_kernel");
  return MICA_OK;
}

extern "C" int mica_cand_neighbor_lists(const double* dis, int64_t m, double max_dis, int cap, int32_t* idx,
                                        int32_t* count, mica_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (m <= 0) return MICA_OK;
  MICA_REQUIRE(dis && idx && count && cap > 0, "bad arguments");
  neighbor_lists_kernel<<<grid_for(m * 32, 256), 256, 0, st>>>(dis, m, max_dis, cap, idx, count);
  MICA_LAUNCH_CHECK("neighbor_lists_kernel");
  return MICA_OK;
}
