// K1: row L2 normalisation forward / backward (HBM-bound, 128-bit vectorised, one warp per row).
// Replaces F.normalize(features, dim=1) -- reference src/models/simclr.py:62,138,294 -- and its autograd
// backward.  Algorithmic bytes: fwd M*D*(b_in + b_out) + 4M ; bwd M*D*(4*n_partials + b_z + b_dp) + 4M.
#include <stdlib.h>

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
  pdl_wait();        // PDL: the preceding kernel of the fused step has completed (no-op otherwise)
  pdl_launch();
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
  pdl_wait();        // PDL: the preceding kernel of the fused step has completed (no-op otherwise)
  pdl_launch();
  constexpr int R = 2;
  const int lane = threadIdx.x & 31;
  const int64_t rows = rows_a + rows_b;
  const bool have = lane < D / kChunk;
  const int64_t row0 = ((int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5)) * R;
  float gv[R][kChunk], zv[R][kChunk], g1[R][kChunk];
  // the second partial slab (the common case: two column splits) is requested together with the first one and the rows
  // of z: one memory round trip instead of two on the launch-bound sizes
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int64_t row = row0 + r;
    if (row < rows && have) {
      load8<float>(dz + row * D + lane * kChunk, gv[r]);
      load8<TZ>(z + row * D + lane * kChunk, zv[r]);
      if (n_partials > 1) load8<float>(dz + partial_stride + row * D + lane * kChunk, g1[r]);
    } else {
#pragma unroll
      for (int i = 0; i < kChunk; ++i) { gv[r][i] = 0.f; zv[r][i] = 0.f; }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int64_t row = row0 + r;
    if (row < rows && have) {
      if (n_partials > 1) {
#pragma unroll
        for (int i = 0; i < kChunk; ++i) gv[r][i] += g1[r][i];
      }
      for (int k = 2; k < n_partials; ++k) {
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

// ---------------- D <= 256, persistent: grid-stride over 4-row groups, the NEXT group's 16-byte loads are issued
// (raw, still packed) before the current group is reduced and stored, so every warp keeps 64 B per lane in flight for
// its whole life instead of only until its first reduction, and no time is lost to block turnover.  kStream adds
// evict-first cache hints (ld.global.cs / st.global.cs) for inputs/outputs far larger than L2. ----------------
template <typename T> struct Raw8 {
  static constexpr int NV = (int)sizeof(T) * 8 / 16;      // 16-byte vectors per 8 elements: 1 (16-bit) or 2 (fp32)
  uint4 v[NV];
};
template <typename T, bool kStream>
__device__ __forceinline__ void raw_load8(const T* p, Raw8<T>& r) {
#pragma unroll
  for (int i = 0; i < Raw8<T>::NV; ++i)
    r.v[i] = kStream ? __ldcs(reinterpret_cast<const uint4*>(p) + i) : __ldg(reinterpret_cast<const uint4*>(p) + i);
}
template <typename T>
__device__ __forceinline__ void raw_unpack8(const Raw8<T>& r, float (&o)[8]) {
  if constexpr (sizeof(T) == 4) {
    o[0] = __uint_as_float(r.v[0].x); o[1] = __uint_as_float(r.v[0].y); o[2] = __uint_as_float(r.v[0].z); o[3] = __uint_as_float(r.v[0].w);
    o[4] = __uint_as_float(r.v[1].x); o[5] = __uint_as_float(r.v[1].y); o[6] = __uint_as_float(r.v[1].z); o[7] = __uint_as_float(r.v[1].w);
  } else {
    const uint32_t w[4] = {r.v[0].x, r.v[0].y, r.v[0].z, r.v[0].w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if constexpr (sizeof(T) == 2 && !__is_same(T, __half)) {          // bf16: the fp32 bit pattern is the high half
        o[2 * q] = __uint_as_float(w[q] << 16); o[2 * q + 1] = __uint_as_float(w[q] & 0xFFFF0000u);
      } else {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[q]));
        o[2 * q] = f.x; o[2 * q + 1] = f.y;
      }
    }
  }
}
template <typename T, bool kStream>
__device__ __forceinline__ void store8s(T* p, const float (&o)[8]) {
  if constexpr (!kStream) {
    store8<T>(p, o);
  } else if constexpr (sizeof(T) == 4) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(o[0], o[1], o[2], o[3]));
    __stcs(reinterpret_cast<float4*>(p) + 1, make_float4(o[4], o[5], o[6], o[7]));
  } else {
    uint4 v;
    if constexpr (__is_same(T, __half)) {
      __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
      for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(o[2 * i], o[2 * i + 1]);
    } else {
      __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
      for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
    }
    __stcs(reinterpret_cast<uint4*>(p), v);
  }
}

template <typename TIn> struct PersistRows { static constexpr int R = sizeof(TIn) == 4 ? 2 : 4; };

template <typename TIn, typename TOut, bool kStream>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
l2norm_fwd_persist_kernel(const TIn* __restrict__ pa, int64_t rows_a, const TIn* __restrict__ pb, int64_t rows_b, int D,
                          TOut* __restrict__ z, float* __restrict__ inv_norm, float eps) {
  pdl_wait();        // PDL: the preceding kernel of the fused step has completed (no-op otherwise)
  pdl_launch();
  const int lane = threadIdx.x & 31;
  const int64_t rows = rows_a + rows_b;
  const bool have = lane < D / kChunk;
  constexpr int kRows = PersistRows<TIn>::R;           // 64 B per lane in flight either way
  const int64_t n_groups = (rows + kRows - 1) / kRows;
  const int64_t gstride = (int64_t)gridDim.x * kWarpsPerBlock;
  int64_t grp = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  Raw8<TIn> cur[kRows], nxt[kRows];
  auto issue = [&](int64_t g, Raw8<TIn> (&dst)[kRows]) {
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const int64_t row = g * kRows + r;
      if (row < rows && have) {
        const TIn* src = row < rows_a ? pa + row * D : pb + (row - rows_a) * D;
        raw_load8<TIn, kStream>(src + lane * kChunk, dst[r]);
      } else {
#pragma unroll
        for (int i = 0; i < Raw8<TIn>::NV; ++i) dst[r].v[i] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  };
  if (grp < n_groups) issue(grp, cur);
  for (; grp < n_groups; grp += gstride) {
    if (grp + gstride < n_groups) issue(grp + gstride, nxt);
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const int64_t row = grp * kRows + r;
      float v[kChunk];
      raw_unpack8<TIn>(cur[r], v);
      float ss = 0.f;
#pragma unroll
      for (int i = 0; i < kChunk; ++i) ss = fmaf(v[i], v[i], ss);
      ss = warp_sum(ss);
      const float inv = 1.0f / fmaxf(sqrtf(ss), eps);
      if (row < rows) {
        if (have) {
#pragma unroll
          for (int i = 0; i < kChunk; ++i) v[i] *= inv;
          store8s<TOut, kStream>(z + row * D + lane * kChunk, v);
        }
        if (lane == 0) inv_norm[row] = inv;
      }
    }
#pragma unroll
    for (int r = 0; r < kRows; ++r) cur[r] = nxt[r];
  }
}

// Experimental (SM3_K1_BWD_VARIANT=1, 16-bit z, D <= 256): the backward in the same persistent form -- a warp owns
// 2-row groups and requests the next group's first gradient slab (fp32, two 16-byte vectors per lane and row) and z row
// (one vector) before it reduces the current group.  Further slabs (n_partials > 1) are read in the loop as before.
template <typename TZ, typename TOut>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
l2norm_bwd_persist_kernel(const float* __restrict__ dz, int n_partials, int64_t partial_stride, float scale,
                          const TZ* __restrict__ z, const float* __restrict__ inv_norm, float inv_eps,
                          TOut* __restrict__ dpa, int64_t rows_a, TOut* __restrict__ dpb, int64_t rows_b, int D) {
  pdl_wait();        // PDL: the preceding kernel of the fused step has completed (no-op otherwise)
  pdl_launch();
  static_assert(sizeof(TZ) == 2, "16-bit z rows");
  constexpr int R = 2;
  const int lane = threadIdx.x & 31;
  const int64_t rows = rows_a + rows_b;
  const bool have = lane < D / kChunk;
  const int64_t n_groups = (rows + R - 1) / R;
  const int64_t gstride = (int64_t)gridDim.x * kWarpsPerBlock;
  int64_t grp = (int64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
  Raw8<float> cg[R], ng[R];
  Raw8<TZ> cz[R], nz[R];
  auto issue = [&](int64_t g, Raw8<float> (&dg)[R], Raw8<TZ> (&dzr)[R]) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t row = g * R + r;
      if (row < rows && have) {
        raw_load8<float, false>(dz + row * D + lane * kChunk, dg[r]);
        raw_load8<TZ, false>(z + row * D + lane * kChunk, dzr[r]);
      } else {
        dg[r].v[0] = dg[r].v[1] = make_uint4(0u, 0u, 0u, 0u);
        dzr[r].v[0] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
  };
  if (grp < n_groups) issue(grp, cg, cz);
  for (; grp < n_groups; grp += gstride) {
    if (grp + gstride < n_groups) issue(grp + gstride, ng, nz);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t row = grp * R + r;
      float gv[kChunk], zv[kChunk];
      raw_unpack8<float>(cg[r], gv);
      raw_unpack8<TZ>(cz[r], zv);
      if (row < rows && have) {
        for (int k = 1; k < n_partials; ++k) {
          float t[kChunk];
          load8<float>(dz + k * partial_stride + row * D + lane * kChunk, t);
#pragma unroll
          for (int i = 0; i < kChunk; ++i) gv[i] += t[i];
        }
      }
      float dot = 0.f;
#pragma unroll
      for (int i = 0; i < kChunk; ++i) { gv[i] *= scale; dot = fmaf(gv[i], zv[i], dot); }
      dot = warp_sum(dot);
      if (row < rows && have) {
        const float inv = inv_norm[row];
        if (inv >= inv_eps) dot = 0.f;          // ||p|| <= eps: F.normalize divides by the constant eps
        TOut* dst = row < rows_a ? dpa + row * D : dpb + (row - rows_a) * D;
#pragma unroll
        for (int i = 0; i < kChunk; ++i) gv[i] = (gv[i] - zv[i] * dot) * inv;
        store8<TOut>(dst + lane * kChunk, gv);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) { cg[r] = ng[r]; cz[r] = nz[r]; }
  }
}

// ---------------- forward ----------------
template <typename TIn, typename TOut, bool kVec>
__global__ void __launch_bounds__(kWarpsPerBlock * 32)
l2norm_fwd_kernel(const TIn* __restrict__ pa, int64_t rows_a, const TIn* __restrict__ pb, int64_t rows_b, int D,
                  TOut* __restrict__ z, float* __restrict__ inv_norm, float eps) {
  pdl_wait();        // PDL: the preceding kernel of the fused step has completed (no-op otherwise)
  pdl_launch();
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
  pdl_wait();        // PDL: the preceding kernel of the fused step has completed (no-op otherwise)
  pdl_launch();
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

// SM3_K1_FWD_VARIANT: 0 = one short-lived CTA per 32 rows, 1 = persistent + register prefetch, 2 = 1 with evict-first
// hints.  Unset = by shape, from the measurements on B200 at 2M x 256 (tools/hbm_variants.py, fraction of the 6.5 TB/s
// copy rate): 16-bit rows 0.71 / 0.85 / 0.88 for variants 0 / 1 / 2 (and 20 -> 16 us at the L2-resident 65536 x 256),
// fp32 rows 1.08 / 0.93 / 0.95 (the read-heavy form is already past the copy rate).  The evict-first hints are only used
// when input + output cannot stay in the 126 MB L2 anyway, so that K2 still finds a small z there.
int k1_fwd_variant(int64_t rows, int D, int in_bytes, int out_bytes) {
  const char* e = getenv("SM3_K1_FWD_VARIANT");
  if (e && e[0] >= '0' && e[0] <= '2') return e[0] - '0';
  if (in_bytes == 4) return 0;
  return rows * (int64_t)D * (in_bytes + out_bytes) > (int64_t)96 << 20 ? 2 : 1;
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
    const int variant = k1_fwd_variant(rows, D, dtype_size(p_dtype), dtype_size(z_dtype));
    if (variant != 0) {
      // persistent grid: exactly as many CTAs as are resident at once (occupancy query, cached per instantiation)
      const int rpw = dtype_size(p_dtype) == 4 ? 2 : 4;                        // PersistRows<TIn>::R
      const int64_t wantp = (rows + kWarpsPerBlock * rpw - 1) / (kWarpsPerBlock * rpw);
      SM3_DISPATCH_DTYPE(p_dtype, TIn, SM3_DISPATCH_DTYPE(z_dtype, TOut, {
        if (variant == 2) {
          static const int per_sm = resident_ctas(l2norm_fwd_persist_kernel<TIn, TOut, true>, kWarpsPerBlock * 32);
          const int64_t cap = (int64_t)num_sms() * per_sm;
          launch_k(l2norm_fwd_persist_kernel<TIn, TOut, true>, dim3((unsigned)(wantp < cap ? wantp : cap)), dim3(kWarpsPerBlock * 32), 0, st, 
              (const TIn*)p_a, rows_a, (const TIn*)p_b, rows_b, D, (TOut*)z, inv_norm, eps);
        } else {
          static const int per_sm = resident_ctas(l2norm_fwd_persist_kernel<TIn, TOut, false>, kWarpsPerBlock * 32);
          const int64_t cap = (int64_t)num_sms() * per_sm;
          launch_k(l2norm_fwd_persist_kernel<TIn, TOut, false>, dim3((unsigned)(wantp < cap ? wantp : cap)), dim3(kWarpsPerBlock * 32), 0, st, 
              (const TIn*)p_a, rows_a, (const TIn*)p_b, rows_b, D, (TOut*)z, inv_norm, eps);
        }
      }));
      SM3_CHECK_CUDA(cudaGetLastError());
      return SM3_OK;
    }
    SM3_DISPATCH_DTYPE(p_dtype, TIn, SM3_DISPATCH_DTYPE(z_dtype, TOut, {
      launch_k(l2norm_fwd_small_kernel<TIn, TOut>, dim3(g4), dim3(kWarpsPerBlock * 32), 0, st, 
          (const TIn*)p_a, rows_a, (const TIn*)p_b, rows_b, D, (TOut*)z, inv_norm, eps);
    }));
    SM3_CHECK_CUDA(cudaGetLastError());
    return SM3_OK;
  }
  SM3_DISPATCH_DTYPE(p_dtype, TIn, SM3_DISPATCH_DTYPE(z_dtype, TOut, {
    if (vec)
      launch_k(l2norm_fwd_kernel<TIn, TOut, true>, dim3(grid), dim3(kWarpsPerBlock * 32), 0, st, 
          (const TIn*)p_a, rows_a, (const TIn*)p_b, rows_b, D, (TOut*)z, inv_norm, eps);
    else
      launch_k(l2norm_fwd_kernel<TIn, TOut, false>, dim3(grid), dim3(kWarpsPerBlock * 32), 0, st, 
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
    const char* ev = getenv("SM3_K1_BWD_VARIANT");           // 1 = experimental persistent + prefetch form (opt-in)
    if (ev && ev[0] == '1' && z_dtype != SM3_F32) {
#define SM3_K1_BWD_PERSIST(TZ, TOut)                                                                                 \
      do {                                                                                                           \
        static const int per_sm = resident_ctas(l2norm_bwd_persist_kernel<TZ, TOut>, kWarpsPerBlock * 32);           \
        const int64_t cap = (int64_t)num_sms() * per_sm;                                                             \
        launch_k(l2norm_bwd_persist_kernel<TZ, TOut>, dim3((unsigned)(want2 < cap ? want2 : cap)), dim3(kWarpsPerBlock * 32), 0, st,   \
            dz_partials, n_partials, rows * (int64_t)D, scale, (const TZ*)z, inv_norm, inv_eps, (TOut*)dp_a, rows_a, \
            (TOut*)dp_b, rows_b, D);                                                                                 \
      } while (0)
      if (z_dtype == SM3_BF16) { SM3_DISPATCH_DTYPE(dp_dtype, TOut, { SM3_K1_BWD_PERSIST(__nv_bfloat16, TOut); }); }
      else { SM3_DISPATCH_DTYPE(dp_dtype, TOut, { SM3_K1_BWD_PERSIST(__half, TOut); }); }
#undef SM3_K1_BWD_PERSIST
      SM3_CHECK_CUDA(cudaGetLastError());
      return SM3_OK;
    }
    SM3_DISPATCH_DTYPE(z_dtype, TZ, SM3_DISPATCH_DTYPE(dp_dtype, TOut, {
      launch_k(l2norm_bwd_small_kernel<TZ, TOut>, dim3(g2), dim3(kWarpsPerBlock * 32), 0, st, 
          dz_partials, n_partials, rows * (int64_t)D, scale, (const TZ*)z, inv_norm, inv_eps, (TOut*)dp_a, rows_a,
          (TOut*)dp_b, rows_b, D);
    }));
    SM3_CHECK_CUDA(cudaGetLastError());
    return SM3_OK;
  }
  SM3_DISPATCH_DTYPE(z_dtype, TZ, SM3_DISPATCH_DTYPE(dp_dtype, TOut, {
    if (vec)
      launch_k(l2norm_bwd_kernel<TZ, TOut, true>, dim3(grid), dim3(kWarpsPerBlock * 32), 0, st, 
          dz_partials, n_partials, rows * (int64_t)D, scale, (const TZ*)z, inv_norm, inv_eps, (TOut*)dp_a, rows_a,
          (TOut*)dp_b, rows_b, D);
    else
      launch_k(l2norm_bwd_kernel<TZ, TOut, false>, dim3(grid), dim3(kWarpsPerBlock * 32), 0, st, 
          dz_partials, n_partials, rows * (int64_t)D, scale, (const TZ*)z, inv_norm, inv_eps, (TOut*)dp_a, rows_a,
          (TOut*)dp_b, rows_b, D);
  }));
  SM3_CHECK_CUDA(cudaGetLastError());
  return SM3_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// upstream scaling of eagerly computed gradients: out_t = in_t * (*g) for up to 8 same-sized tensors in ONE launch
// (autograd hands the fused-step Functions a device scalar: GradScaler factor x loss weight, backbone_train.py:101-125).
// ---------------------------------------------------------------------------------------------------------------
namespace {
struct ScaleArgs {
  const void* in[8];
  void* out[8];
  int count;
};
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256)
scale_grads_kernel(ScaleArgs a, int64_t n, const float* __restrict__ g) {
  const float s = __ldg(g);
  const TIn* src = reinterpret_cast<const TIn*>(a.in[blockIdx.y]);
  TOut* dst = reinterpret_cast<TOut*>(a.out[blockIdx.y]);
  const int64_t n8 = n / 8;
  const bool vec = (((uintptr_t)src | (uintptr_t)dst) & 31u) == 0;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec) {
    for (int64_t i = t0; i < n8; i += stride) {
      float v[8];
      load8<TIn>(src + i * 8, v);
#pragma unroll
      for (int k = 0; k < 8; ++k) v[k] *= s;
      store8<TOut>(dst + i * 8, v);
    }
    for (int64_t j = n8 * 8 + t0; j < n; j += stride) dst[j] = from_f32<TOut>(to_f32(src[j]) * s);
  } else {
    for (int64_t i = t0; i < n; i += stride) dst[i] = from_f32<TOut>(to_f32(src[i]) * s);
  }
}
}  // namespace

}  // namespace sm3

extern "C" int sm3_scale_grads(const void* const* in_host_array, void* const* out_host_array, int count, int64_t numel,
                               int in_dtype, int out_dtype, const float* g_device, void* stream) {
  using namespace sm3;
  SM3_REQUIRE(in_host_array && out_host_array && g_device, SM3_ERR_SHAPE, "scale_grads: null pointer");
  SM3_REQUIRE(count >= 1 && count <= 8 && numel >= 0, SM3_ERR_SHAPE, "scale_grads: count=%d not in [1,8]", count);
  SM3_REQUIRE(dtype_ok(in_dtype) && dtype_ok(out_dtype), SM3_ERR_DTYPE, "scale_grads: bad dtype");
  if (numel == 0) return SM3_OK;
  ScaleArgs a{};
  a.count = count;
  for (int t = 0; t < count; ++t) {
    SM3_REQUIRE(in_host_array[t] && out_host_array[t], SM3_ERR_SHAPE, "scale_grads: tensor %d is null", t);
    a.in[t] = in_host_array[t];
    a.out[t] = out_host_array[t];
  }
  int64_t blocks = (numel / 8 + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 8 / count + 1;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  SM3_DISPATCH_DTYPE(in_dtype, TIn, SM3_DISPATCH_DTYPE(out_dtype, TOut, {
    scale_grads_kernel<TIn, TOut><<<dim3((unsigned)blocks, (unsigned)count), 256, 0, (cudaStream_t)stream>>>(a, numel, g_device);
  }));
  SM3_CHECK_CUDA(cudaGetLastError());
  return SM3_OK;
}
