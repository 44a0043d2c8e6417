// R1: cubic B-spline (and trilinear) resampling of the cryo-EM map.
//
// Replaces scipy.ndimage.zoom(data, zoom, order=3) as called at
// utils/preprocessing.py:117 (reference root) -- SciPy's algorithm restated in
// float64 IIR prefilter with mirror boundaries along each
// axis, align-corners coordinate map, 4x4x4 tap gather with mirrored edge taps,
// float32 result.  SciPy quirk kept on purpose: an output index
// whose coordinate k*(n_in-1)/(n_out-1) overshoots n_in-1 by rounding is treated
// as outside the map (mode='constant') and yields 0.
//
// Kernels (all HBM-bound, see DESIGN.md for the byte counts):
//   taps_kernel          per-axis tap index / weight tables (tiny)
//   prefilter_cols       IIR along a strided axis (z or y): a [n x CW] float64 tile
//                        of CW neighbouring lines is staged in shared memory with
//                        coalesced loads, one thread sweeps each line, coalesced store
//   prefilter_rows       IIR along the contiguous axis (x): [R x n] tile, same idea
//   gather3 / gather1    64-tap / 8-tap separable gather, x fastest across threads
#include "common.cuh"

namespace mica {

struct Tap {
  int idx[4];
  double w[4];
};

__device__ __forceinline__ int mirror_index(long idx, int len) {
  // SciPy ni_interpolation.c edge handling for NI_EXTEND_MIRROR-like taps
  if (len <= 1) return 0;
  long s2 = 2L * len - 2;
  if (idx < 0) {
    idx = s2 * (long)(-idx / s2) + idx;
    idx = (idx <= 1 - len) ? idx + s2 : -idx;
  } else if (idx >= len) {
    idx -= s2 * (long)(idx / s2);
    if (idx >= len) idx = s2 - idx;
  }
  return (int)idx;
}

__global__ void taps_kernel(Tap* __restrict__ taps, int n_local, int k0, int n_in, int n_out, int order,
                            int in_off, int in_local) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_local) return;
  int k = k0 + i;
  double zoom = (n_out > 1) ? (double)(n_in - 1) / (double)(n_out - 1) : 1.0;
  double cc = (double)k * zoom;
  Tap t;
  if (cc < 0.0 || cc > (double)(n_in - 1)) {  // SciPy map_coordinate(NI_EXTEND_CONSTANT) -> cval
    for (int l = 0; l < 4; ++l) {
      t.idx[l] = 0;
      t.w[l] = 0.0;
    }
  } else {
    double fl = floor(cc);
    double x = cc - fl;
    long start;
    int ntap;
    if (order == 3) {
      start = (long)fl - 1;
      ntap = 4;
      double y = x, z = 1.0 - x;
      t.w[1] = (y * y * (y - 2.0) * 3.0 + 4.0) / 6.0;
      t.w[2] = (z * z * (z - 2.0) * 3.0 + 4.0) / 6.0;
      t.w[0] = z * z * z / 6.0;
      t.w[3] = 1.0 - t.w[0] - t.w[1] - t.w[2];
    } else {
      start = (long)fl;
      ntap = 2;
      t.w[0] = 1.0 - x;
      t.w[1] = x;
      t.w[2] = t.w[3] = 0.0;
    }
    for (int l = 0; l < 4; ++l) {
      int g = (l < ntap) ? mirror_index(start + l, n_in) : 0;
      int loc = g - in_off;
      loc = loc < 0 ? 0 : (loc >= in_local ? in_local - 1 : loc);
      t.idx[l] = (l < ntap) ? loc : 0;
    }
  }
  taps[i] = t;
}

// ---------------------------------------------------------------- prefilter
constexpr double kPole = -0.26794919243112270647;  // sqrt(3) - 2
constexpr int kInitHorizon = 56;                    // |pole|^56 ~ 1e-32: below double rounding

// One thread filters one line held in shared memory; `stride` is the distance (in
// doubles) between successive samples of the line inside the tile.
__device__ __forceinline__ void iir_line(double* line, int n, int stride) {
  if (n < 2) return;
  const double z = kPole;
  // causal initialisation, mirror boundary, SciPy's pairing of k and n-1-k
  double z_n_1 = pow(z, (double)(n - 1));
  double c0 = line[0] + z_n_1 * line[(n - 1) * stride];
  double z_i = z;
  int lim = min(n - 1, kInitHorizon);
  for (int k = 1; k < lim; ++k) {
    c0 += z_i * (line[k * stride] + z_n_1 * line[(n - 1 - k) * stride]);
    z_i *= z;
  }
  c0 /= (1.0 - z_n_1 * z_n_1);
  line[0] = c0;
  double prev = c0;
#pragma unroll 4
  for (int k = 1; k < n; ++k) {
    prev = fma(z, prev, line[k * stride]);
    line[k * stride] = prev;
  }
  // anticausal
  double last = (z / (z * z - 1.0)) * (prev + z * line[(n - 2) * stride]);
  line[(n - 1) * stride] = last;
  prev = last;
#pragma unroll 4
  for (int k = n - 2; k >= 0; --k) {
    prev = z * (prev - line[k * stride]);
    line[k * stride] = prev;
  }
}

constexpr double kGain = (1.0 - kPole) * (1.0 - 1.0 / kPole);  // = 6

// lines along a strided axis.  grid = (ceil(n_cols / CW), n_outer)
template <typename TIn, int CW>
__global__ void __launch_bounds__(256)
prefilter_cols(const TIn* __restrict__ in, double* __restrict__ out, int n, int64_t line_stride,
               int n_cols, int64_t outer_stride) {
  extern __shared__ double tile[];  // [n][CW]
  const int col0 = blockIdx.x * CW;
  const int64_t base = (int64_t)blockIdx.y * outer_stride + col0;
  const int ncol = min(CW, n_cols - col0);
  const double gain = (n >= 2) ? kGain : 1.0;  // SciPy leaves length-1 axes unfiltered
  for (int e = threadIdx.x; e < n * CW; e += blockDim.x) {
    int k = e / CW, j = e % CW;
    if (j < ncol) tile[e] = (double)in[base + (int64_t)k * line_stride + j] * gain;
  }
  __syncthreads();
  if (threadIdx.x < ncol) iir_line(tile + threadIdx.x, n, CW);
  __syncthreads();
  for (int e = threadIdx.x; e < n * CW; e += blockDim.x) {
    int k = e / CW, j = e % CW;
    if (j < ncol) out[base + (int64_t)k * line_stride + j] = tile[e];
  }
}

// lines along the contiguous axis.  grid = ceil(n_rows / R); tile [R][n + 1]
template <int R>
__global__ void __launch_bounds__(256)
prefilter_rows(double* __restrict__ data, int n, int64_t n_rows) {
  extern __shared__ double tile[];
  const int pitch = n | 1;  // odd pitch: rows start in different banks
  const int64_t row0 = (int64_t)blockIdx.x * R;
  const int nrow = (int)min((int64_t)R, n_rows - row0);
  double* g = data + row0 * n;
  const double gain = (n >= 2) ? kGain : 1.0;
  for (int e = threadIdx.x; e < nrow * n; e += blockDim.x) {
    int r = e / n, k = e - r * n;
    tile[r * pitch + k] = g[e] * gain;
  }
  __syncthreads();
  if (threadIdx.x < nrow) iir_line(tile + threadIdx.x * pitch, n, 1);
  __syncthreads();
  for (int e = threadIdx.x; e < nrow * n; e += blockDim.x) {
    int r = e / n, k = e - r * n;
    g[e] = tile[r * pitch + k];
  }
}

// fallback for lines too long for shared memory: sweep in global memory
template <typename TIn>
__global__ void prefilter_cols_global(const TIn* __restrict__ in, double* __restrict__ out, int n,
                                      int64_t line_stride, int n_cols, int64_t outer_stride) {
  int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= n_cols) return;
  const int64_t base = (int64_t)blockIdx.y * outer_stride + col;
  const double gain = (n >= 2) ? kGain : 1.0;
  for (int k = 0; k < n; ++k) out[base + k * line_stride] = (double)in[base + k * line_stride] * gain;
  // iir_line with a 64-bit stride
  double* line = out + base;
  const double z = kPole;
  if (n < 2) return;
  double z_n_1 = pow(z, (double)(n - 1));
  double c0 = line[0] + z_n_1 * line[(n - 1) * line_stride];
  double z_i = z;
  int lim = min(n - 1, kInitHorizon);
  for (int k = 1; k < lim; ++k) {
    c0 += z_i * (line[k * line_stride] + z_n_1 * line[(n - 1 - k) * line_stride]);
    z_i *= z;
  }
  c0 /= (1.0 - z_n_1 * z_n_1);
  line[0] = c0;
  double prev = c0;
  for (int k = 1; k < n; ++k) {
    prev = fma(z, prev, line[k * line_stride]);
    line[k * line_stride] = prev;
  }
  double last = (z / (z * z - 1.0)) * (prev + z * line[(n - 2) * line_stride]);
  line[(n - 1) * line_stride] = last;
  prev = last;
  for (int k = n - 2; k >= 0; --k) {
    prev = z * (prev - line[k * line_stride]);
    line[k * line_stride] = prev;
  }
}

// ------------------------------------------------------------------- gather
// grid = (ceil(nx/128), ny, nz_local); one output voxel per thread
__global__ void __launch_bounds__(128)
gather3_kernel(const double* __restrict__ c, int sy, int sx, const Tap* __restrict__ tz,
               const Tap* __restrict__ ty, const Tap* __restrict__ tx, float* __restrict__ dst, int ny,
               int nx) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= nx) return;
  const int y = blockIdx.y, z = blockIdx.z;
  const Tap Tx = tx[x];
  const Tap Ty = ty[y];
  const Tap Tz = tz[z];
  const int64_t plane = (int64_t)sy * sx;
  double acc = 0.0;
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    const double* pz = c + (int64_t)Tz.idx[l] * plane;
    double accy = 0.0;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      const double* row = pz + (int64_t)Ty.idx[m] * sx;
      double r = Tx.w[0] * row[Tx.idx[0]] + Tx.w[1] * row[Tx.idx[1]] + Tx.w[2] * row[Tx.idx[2]] +
                 Tx.w[3] * row[Tx.idx[3]];
      accy += Ty.w[m] * r;
    }
    acc += Tz.w[l] * accy;
  }
  st_stream(dst + ((int64_t)z * ny + y) * nx + x, (float)acc);
}

__global__ void __launch_bounds__(128)
gather1_kernel(const float* __restrict__ src, int sy, int sx, const Tap* __restrict__ tz,
               const Tap* __restrict__ ty, const Tap* __restrict__ tx, float* __restrict__ dst, int ny,
               int nx) {
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= nx) return;
  const int y = blockIdx.y, z = blockIdx.z;
  const Tap Tx = tx[x];
  const Tap Ty = ty[y];
  const Tap Tz = tz[z];
  const int64_t plane = (int64_t)sy * sx;
  double acc = 0.0;
#pragma unroll
  for (int l = 0; l < 2; ++l) {
    const float* pz = src + (int64_t)Tz.idx[l] * plane;
    double accy = 0.0;
#pragma unroll
    for (int m = 0; m < 2; ++m) {
      const float* row = pz + (int64_t)Ty.idx[m] * sx;
      double r = Tx.w[0] * (double)row[Tx.idx[0]] + Tx.w[1] * (double)row[Tx.idx[1]];
      accy += Ty.w[m] * r;
    }
    acc += Tz.w[l] * accy;
  }
  st_stream(dst + ((int64_t)z * ny + y) * nx + x, (float)acc);
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

constexpr size_t kMaxTileBytes = 200 * 1024;

template <typename TIn>
static int launch_cols(const TIn* in, double* out, int n, int64_t line_stride, int n_cols,
                       int64_t outer_stride, int n_outer, cudaStream_t st) {
  if (n_cols <= 0 || n_outer <= 0 || n <= 0) return MICA_OK;
  size_t per_col = (size_t)n * sizeof(double);
  if (per_col * 32 <= kMaxTileBytes) {
    size_t smem = per_col * 32;
    MICA_CUDA(cudaFuncSetAttribute(prefilter_cols<TIn, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    prefilter_cols<TIn, 32><<<dim3((n_cols + 31) / 32, n_outer), 256, smem, st>>>(in, out, n, line_stride, n_cols, outer_stride);
  } else if (per_col * 8 <= kMaxTileBytes) {
    size_t smem = per_col * 8;
    MICA_CUDA(cudaFuncSetAttribute(prefilter_cols<TIn, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    prefilter_cols<TIn, 8><<<dim3((n_cols + 7) / 8, n_outer), 256, smem, st>>>(in, out, n, line_stride, n_cols, outer_stride);
  } else {
    prefilter_cols_global<TIn><<<dim3((n_cols + 127) / 128, n_outer), 128, 0, st>>>(in, out, n, line_stride, n_cols, outer_stride);
  }
  MICA_LAUNCH_CHECK("prefilter_cols");
  return MICA_OK;
}

static int launch_rows(double* data, int n, int64_t n_rows, cudaStream_t st) {
  if (n <= 0 || n_rows <= 0) return MICA_OK;
  size_t per_row = (size_t)(n | 1) * sizeof(double);
  if (per_row * 32 <= kMaxTileBytes) {
    size_t smem = per_row * 32;
    MICA_CUDA(cudaFuncSetAttribute(prefilter_rows<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    prefilter_rows<32><<<(unsigned)ceil_div64(n_rows, 32), 256, smem, st>>>(data, n, n_rows);
  } else if (per_row * 4 <= kMaxTileBytes) {
    size_t smem = per_row * 4;
    MICA_CUDA(cudaFuncSetAttribute(prefilter_rows<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    prefilter_rows<4><<<(unsigned)ceil_div64(n_rows, 4), 256, smem, st>>>(data, n, n_rows);
  } else {
    // a row is a column of a [n_rows x n] matrix with unit line stride
    prefilter_cols_global<double><<<dim3(1, (unsigned)n_rows), 1, 0, st>>>(data, data, n, 1, 1, n);
  }
  MICA_LAUNCH_CHECK("prefilter_rows");
  return MICA_OK;
}

}  // namespace mica

using namespace mica;

extern "C" int mica_zoom_output_shape(const int in_zyx[3], const float zoom_zyx[3], int out_zyx[3]) {
  MICA_REQUIRE(in_zyx && zoom_zyx && out_zyx, "null argument");
  for (int a = 0; a < 3; ++a) {
    volatile float prod = (float)in_zyx[a] * zoom_zyx[a];  // float32 product (NumPy 2 weak-scalar promotion)
    out_zyx[a] = (int)nearbyint((double)prod);              // round-half-even, as python round()
  }
  return MICA_OK;
}

extern "C" size_t mica_resample_workspace_bytes(int src_nz_local, int sy, int sx, int nz, int ny, int nx, int order) {
  size_t taps = align_up((size_t)(nz + ny + nx) * sizeof(Tap), 256);
  size_t coeff = (order == 3) ? align_up((size_t)src_nz_local * sy * sx * sizeof(double), 256) : 0;
  return taps + coeff + 256;
}

extern "C" int mica_bspline_resample_f32(const float* src, int sz, int sy, int sx, int src_z0, int src_nz_local,
                                         float* dst, int nz, int ny, int nx, int dst_z0, int dst_nz_local,
                                         void* workspace, size_t workspace_bytes, int order, mica_stream_t stream) {
  cudaStream_t st = (cudaStream_t)stream;
  MICA_REQUIRE(src && dst && workspace, "null pointer");
  MICA_REQUIRE(order == 3 || order == 1, "order must be 3 or 1 (got %d)", order);
  MICA_REQUIRE(sz > 0 && sy > 0 && sx > 0 && nz > 0 && ny > 0 && nx > 0, "empty shape");
  MICA_REQUIRE(src_z0 >= 0 && src_nz_local > 0 && src_z0 + src_nz_local <= sz, "bad source slab");
  MICA_REQUIRE(dst_z0 >= 0 && dst_nz_local >= 0 && dst_z0 + dst_nz_local <= nz, "bad output slab");
  MICA_REQUIRE(ny <= 65535 && dst_nz_local <= 65535, "output too large for the launch grid");
  if (workspace_bytes < mica_resample_workspace_bytes(src_nz_local, sy, sx, nz, ny, nx, order))
    return set_error(MICA_ERR_WORKSPACE, "resample workspace too small");
  if (dst_nz_local == 0) return MICA_OK;

  char* ws = (char*)(((uintptr_t)workspace + 255) / 256 * 256);
  Tap* tz = (Tap*)ws;
  Tap* ty = tz + nz;
  Tap* tx = ty + ny;
  double* coeff = (double*)(ws + align_up((size_t)(nz + ny + nx) * sizeof(Tap), 256));

  taps_kernel<<<(dst_nz_local + 127) / 128, 128, 0, st>>>(tz, dst_nz_local, dst_z0, sz, nz, order, src_z0, src_nz_local);
  MICA_LAUNCH_CHECK("taps_kernel(z)");
  taps_kernel<<<(ny + 127) / 128, 128, 0, st>>>(ty, ny, 0, sy, ny, order, 0, sy);
  MICA_LAUNCH_CHECK("taps_kernel(y)");
  taps_kernel<<<(nx + 127) / 128, 128, 0, st>>>(tx, nx, 0, sx, nx, order, 0, sx);
  MICA_LAUNCH_CHECK("taps_kernel(x)");

  dim3 grid((nx + 127) / 128, ny, dst_nz_local);
  if (order == 3) {
    const int64_t plane = (int64_t)sy * sx;
    MICA_REQUIRE(plane <= 0x7fffffffLL && src_nz_local <= 65535, "source plane too large");
    // axis 0 (z): lines of length src_nz_local, stride plane; reads float32, writes float64
    int rc = launch_cols<float>(src, coeff, src_nz_local, plane, (int)plane, 0, 1, st);
    if (rc) return rc;
    // axis 1 (y): per z plane, lines of length sy, stride sx
    rc = launch_cols<double>(coeff, coeff, sy, sx, sx, plane, src_nz_local, st);
    if (rc) return rc;
    // axis 2 (x): contiguous rows
    rc = launch_rows(coeff, sx, (int64_t)src_nz_local * sy, st);
    if (rc) return rc;
    gather3_kernel<<<grid, 128, 0, st>>>(coeff, sy, sx, tz, ty, tx, dst, ny, nx);
    MICA_LAUNCH_CHECK("gather3_kernel");
  } else {
    gather1_kernel<<<grid, 128, 0, st>>>(src, sy, sx, tz, ty, tx, dst, ny, nx);
    MICA_LAUNCH_CHECK("gather1_kernel");
  }
  return MICA_OK;
}
