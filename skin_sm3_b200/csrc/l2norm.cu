// K1: row L2 normalisation forward / backward (HBM-bound, 128-bit vectorised, one warp per row).
// Replaces F.normalize(features, dim=1) -- reference src/models/simclr.py:62,138,294 -- and its autograd
// backward.  Algorithmic bytes: fwd M*D*(b_in + b_out) + 4M ; bwd M*D*(4*n_partials + b_z + b_dp) + 4M.
#include "common.cuh"

namespace sm3 {
namespace {

constexpr int kChunk = 8;        // elements per lane per step (one or two 16-byte transactions)
constexpr int kMaxChunks = 4;    // register-cached row: D <= 8 * 32 * 4 = 1024
constexpr int kWarpsPerBlock = 8;

template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&o)[8]) {
  if constexpr (sizeof(T) == 4) {
    float a[4], b[4];
    VecIO<float>::load(reinterpret_cast<const float*>(p), a);
    VecIO<float>::load(reinterpret_cast<const float*>(p) + 4, b);
#pragma unroll
    for (int i = 0; i < 4; ++i) { o[i] = a[i]; o[4 + i] = b[i]; }
  } else {
    VecIO<T>::load(p, o);
  }
}
template <typename T>
__device__ __forceinline__ void store8(T* p, const float (&o)[8]) {
  if constexpr (sizeof(T) == 4) {
    float a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { a[i] = o[i]; b[i] = o[4 + i]; }
    VecIO<float>::store(reinterpret_cast<float*>(p), a);
    VecIO<float>::store(reinterpret_cast<float*>(p) + 4, b);
  } else {
    VecIO<T>::store(p, o);
  }
}

// ---------------- D <= 256: kRows rows per warp, all loads issued before the first reduction ----------------
constexpr int kRows = 4;

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
l2norm_fwd_small_kernel(const TIn* __restrict__ pa, int64_t rows_a, const TIn* __restrict__ pb, int64_t rows_b, int D,
                        TOut* __restrict__ z, float* __restrict__ inv_norm, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t rows = rows_a + rows_b;
  const bool have = lane < D / kChunk;
  const int64_t row0 = ((int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5)) * kRows;
  float v[kRows][kChunk];
#pragma unroll
  for (int r = 0; r < kRows; ++r) {
    const int64_t row = row0 + r;
    if (row < rows && have) {
      const TIn* src = row < rows_a ? pa + row * D : pb + (row - rows_a) * D;
      load8<TIn>(src + lane * kChunk, v[r]);
    } else {
#pragma unroll
      for (int i = 0; i < kChunk; ++i) v[r][i] = 0.f;
    }
  }
#pragma unroll
  for (int r = 0; r < kRows; ++r) {
    const int64_t row = row0 + r;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < kChunk; ++i) ss = fmaf(v[r][i], v[r][i], ss);
    ss = warp_sum(ss);
    const float inv = 1.0f / fmaxf(sqrtf(ss), eps);
    if (row < rows) {
      if (have) {
#pragma unroll
        for (int i = 0; i < kChunk; ++i) v[r][i] *= inv;
        store8<TOut>(z + row * D + lane * kChunk, v[r]);
      }
      if (lane == 0) inv_norm[row] = inv;
    }
  }
}

template <typename TZ, typename TOut>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
l2norm_bwd_small_kernel(const float* __restrict__ dz, int n_partials, int64_t partial_stride, float scale,
                        const TZ* __restrict__ z, const float* __restrict__ inv_norm, float inv_eps,
                        TOut* __restrict__ dpa, int64_t rows_a, TOut* __restrict__ dpb, int64_t rows_b, int D) {
  constexpr int R = 2;
  const int lane = threadIdx.x & 31;
  const int64_t rows = rows_a + rows_b;
  const bool have = lane < D / kChunk;
  const int64_t row0 = ((int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5)) * R;
  float gv[R][kChunk], zv[R][kChunk];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int64_t row = row0 + r;
    if (row < rows && have) {
      load8<float>(dz + row * D + lane * kChunk, gv[r]);
      load8<TZ>(z + row * D + lane * kChunk, zv[r]);
    } else {
#pragma unroll
      for (int i = 0; i < kChunk; ++i) { gv[r][i] = 0.f; zv[r][i] = 0.f; }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int64_t row = row0 + r;
    if (row < rows && have) {
      for (int k = 1; k < n_partials; ++k) {
        float t[kChunk];
        load8<float>(dz + k * partial_stride + row * D + lane * kChunk, t);
#pragma unroll
        for (int i = 0; i < kChunk; ++i) gv[r][i] += t[i];
      }
    }
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < kChunk; ++i) { gv[r][i] *= scale; dot = fmaf(gv[r][i], zv[r][i], dot); }
    dot = warp_sum(dot);
    if (row < rows && have) {
      const float inv = inv_norm[row];
      if (inv >= inv_eps) dot = 0.f;          // ||p|| <= eps: F.normalize divides by the constant eps
      TOut* dst = row < rows_a ? dpa + row * D : dpb + (row - rows_a) * D;
#pragma unroll
      for (int i = 0; i < kChunk; ++i) gv[r][i] = (gv[r][i] - zv[r][i] * dot) * inv;
      store8<TOut>(dst + lane * kChunk, gv[r]);
    }
  }
}

// ---------------- forward ----------------
template <typename TIn, typename TOut, bool kVec>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
l2norm_fwd_kernel(const TIn* __restrict__ pa, int64_t rows_a, const TIn* __restrict__ pb, int64_t rows_b, int D,
                  TOut* __restrict__ z, float* __restrict__ inv_norm, float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= rows_a + rows_b) return;
  const TIn* src = row < rows_a ? pa + row * D : pb + (row - rows_a) * D;
  TOut* dst = z + row * D;
  if constexpr (kVec) {
    float v[kMaxChunks][kChunk];
    const int nchunks = D / kChunk;
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nchunks) {
        load8<TIn>(src + ch * kChunk, v[c]);
#pragma unroll
        for (int i = 0; i < kChunk; ++i) ss = fmaf(v[c][i], v[c][i], ss);
      }
    }
    ss = warp_sum(ss);
    const float inv = 1.0f / fmaxf(sqrtf(ss), eps);
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nchunks) {
#pragma unroll
        for (int i = 0; i < kChunk; ++i) v[c][i] *= inv;
        store8<TOut>(dst + ch * kChunk, v[c]);
      }
    }
    if (lane == 0) inv_norm[row] = inv;
  } else {
    float ss = 0.f;
    for (int i = lane; i < D; i += 32) { const float x = to_f32(src[i]); ss = fmaf(x, x, ss); }
    ss = warp_sum(ss);
    const float inv = 1.0f / fmaxf(sqrtf(ss), eps);
    for (int i = lane; i < D; i += 32) dst[i] = from_f32<TOut>(to_f32(src[i]) * inv);
    if (lane == 0) inv_norm[row] = inv;
  }
}

// ---------------- backward ----------------
template <typename TZ, typename TOut, bool kVec>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
l2norm_bwd_kernel(const float* __restrict__ dz, int n_partials, int64_t partial_stride, float scale,
                  const TZ* __restrict__ z, const float* __restrict__ inv_norm, float inv_eps,
                  TOut* __restrict__ dpa, int64_t rows_a, TOut* __restrict__ dpb, int64_t rows_b, int D) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  if (row >= rows_a + rows_b) return;
  const float* g = dz + row * D;
  const TZ* zr = z + row * D;
  TOut* dst = row < rows_a ? dpa + row * D : dpb + (row - rows_a) * D;
  const float inv = inv_norm[row];
  const bool clamped = inv >= inv_eps;  // ||p|| <= eps: F.normalize divides by the constant eps
  if constexpr (kVec) {
    float gv[kMaxChunks][kChunk], zv[kMaxChunks][kChunk];
    const int nchunks = D / kChunk;
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nchunks) {
        load8<float>(g + ch * kChunk, gv[c]);
        for (int k = 1; k < n_partials; ++k) {
          float t[kChunk];
          load8<float>(g + k * partial_stride + ch * kChunk, t);
#pragma unroll
          for (int i = 0; i < kChunk; ++i) gv[c][i] += t[i];
        }
        load8<TZ>(zr + ch * kChunk, zv[c]);
#pragma unroll
        for (int i = 0; i < kChunk; ++i) { gv[c][i] *= scale; dot = fmaf(gv[c][i], zv[c][i], dot); }
      }
    }
    dot = clamped ? 0.f : warp_sum(dot);
#pragma unroll
    for (int c = 0; c < kMaxChunks; ++c) {
      const int ch = lane + 32 * c;
      if (ch < nchunks) {
#pragma unroll
        for (int i = 0; i < kChunk; ++i) gv[c][i] = (gv[c][i] - zv[c][i] * dot) * inv;
        store8<TOut>(dst + ch * kChunk, gv[c]);
      }
    }
  } else {
    float dot = 0.f;
    for (int i = lane; i < D; i += 32) {
      float a = 0.f;
      for (int k = 0; k < n_partials; ++k) a += g[k * partial_stride + i];
      dot = fmaf(a * scale, to_f32(zr[i]), dot);
    }
    dot = clamped ? 0.f : warp_sum(dot);
    for (int i = lane; i < D; i += 32) {
      float a = 0.f;
      for (int k = 0; k < n_partials; ++k) a += g[k * partial_stride + i];
      dst[i] = from_f32<TOut>((a * scale - to_f32(zr[i]) * dot) * inv);
    }
  }
}

}  // namespace

int l2norm_fwd_launch(const void* p_a, int64_t rows_a, const void* p_b, int64_t rows_b, int D, int p_dtype, void* z,
                      int z_dtype, float* inv_norm, float eps, cudaStream_t st) {
  const int64_t rows = rows_a + rows_b;
  if (rows == 0) return SM3_OK;
  const bool vec = (D % kChunk == 0) && D <= kChunk * 32 * kMaxChunks && aligned16(p_a) && aligned16(z) &&
                   (p_b == nullptr || aligned16(p_b));
  const unsigned grid = (unsigned)((rows + kWarpsPerBlock - 1) / kWarpsPerBlock);
  if (vec && D <= 32 * kChunk) {
    const int64_t want4 = (rows + kWarpsPerBlock * kRows - 1) / (kWarpsPerBlock * kRows);
    const unsigned g4 = (unsigned)(want4 < 0x7fffffff ? want4 : 0x7fffffff);
    SM3_DISPATCH_DTYPE(p_dtype, TIn, SM3_DISPATCH_DTYPE(z_dtype, TOut, {
      l2norm_fwd_small_kernel<TIn, TOut><<<g4, kWarpsPerBlock * 32, 0, st>>>(
          (const TIn*)p_a, rows_a, (const TIn*)p_b, rows_b, D, (TOut*)z, inv_norm, eps);
    }));
    SM3_CHECK_CUDA(cudaGetLastError());
    return SM3_OK;
  }
  SM3_DISPATCH_DTYPE(p_dtype, TIn, SM3_DISPATCH_DTYPE(z_dtype, TOut, {
    if (vec)
      l2norm_fwd_kernel<TIn, TOut, true><<<grid, kWarpsPerBlock * 32, 0, st>>>(
          (const TIn*)p_a, rows_a, (const TIn*)p_b, rows_b, D, (TOut*)z, inv_norm, eps);
    else
      l2norm_fwd_kernel<TIn, TOut, false><<<grid, kWarpsPerBlock * 32, 0, st>>>(
          (const TIn*)p_a, rows_a, (const TIn*)p_b, rows_b, D, (TOut*)z, inv_norm, eps);
  }));
  SM3_CHECK_CUDA(cudaGetLastError());
  return SM3_OK;
}

int l2norm_bwd_launch(const float* dz_partials, int n_partials, float scale, const void* z, int z_dtype,
                      const float* inv_norm, float eps, void* dp_a, int64_t rows_a, void* dp_b, int64_t rows_b, int D,
                      int dp_dtype, cudaStream_t st) {
  const int64_t rows = rows_a + rows_b;
  if (rows == 0) return SM3_OK;
  const bool vec = (D % kChunk == 0) && D <= kChunk * 32 * kMaxChunks && aligned16(dz_partials) && aligned16(z) &&
                   aligned16(dp_a) && (dp_b == nullptr || aligned16(dp_b));
  const unsigned grid = (unsigned)((rows + kWarpsPerBlock - 1) / kWarpsPerBlock);
  const float inv_eps = 1.0f / eps;
  if (vec && D <= 32 * kChunk) {
    const int64_t want2 = (rows + kWarpsPerBlock * 2 - 1) / (kWarpsPerBlock * 2);
    const unsigned g2 = (unsigned)(want2 < 0x7fffffff ? want2 : 0x7fffffff);
    SM3_DISPATCH_DTYPE(z_dtype, TZ, SM3_DISPATCH_DTYPE(dp_dtype, TOut, {
      l2norm_bwd_small_kernel<TZ, TOut><<<g2, kWarpsPerBlock * 32, 0, st>>>(
          dz_partials, n_partials, rows * (int64_t)D, scale, (const TZ*)z, inv_norm, inv_eps, (TOut*)dp_a, rows_a,
          (TOut*)dp_b, rows_b, D);
    }));
    SM3_CHECK_CUDA(cudaGetLastError());
    return SM3_OK;
  }
  SM3_DISPATCH_DTYPE(z_dtype, TZ, SM3_DISPATCH_DTYPE(dp_dtype, TOut, {
    if (vec)
      l2norm_bwd_kernel<TZ, TOut, true><<<grid, kWarpsPerBlock * 32, 0, st>>>(
          dz_partials, n_partials, rows * (int64_t)D, scale, (const TZ*)z, inv_norm, inv_eps, (TOut*)dp_a, rows_a,
          (TOut*)dp_b, rows_b, D);
    else
      l2norm_bwd_kernel<TZ, TOut, false><<<grid, kWarpsPerBlock * 32, 0, st>>>(
          dz_partials, n_partials, rows * (int64_t)D, scale, (const TZ*)z, inv_norm, inv_eps, (TOut*)dp_a, rows_a,
          (TOut*)dp_b, rows_b, D);
  }));
  SM3_CHECK_CUDA(cudaGetLastError());
  return SM3_OK;
}

}  // namespace sm3
