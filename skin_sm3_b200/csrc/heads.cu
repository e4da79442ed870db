// K4 multi-head softmax cross-entropy and K5 BCE-with-logits: forward + backward in one launch each,
// coalesced 128-bit global access staged through shared memory, deterministic last-block reduction.
// Also the two tiny per-row kernels that close the InfoNCE forward (partial-sum finalize, CE-on-statistics).
//
// K4 replaces the reference's 8-head loops: tools/mlc_eval.py:159-162, tools/backbone_eval.py:102-105,
// tools/backbone_train.py:178-181, tools/mlc_train.py:255-261 (+ ignore_index=-100 at :381).
// K5 has no reference counterpart; it mirrors torch's binary_cross_entropy_with_logits.
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace sm3 {
namespace {

__device__ __forceinline__ float ex2f_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float lg2f_approx(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpf_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

constexpr int kMaxHeads = 16;
constexpr int kMaxClassesTotal = 64;
constexpr int kHeadThreads = 256;

struct HeadMeta {
  int H, C;                    // heads, total classes
  int offset[kMaxHeads + 1];   // class offset of each head
  float weight[kMaxHeads];
  float gs[kMaxHeads];         // grad_scale * w_h * inv_T / (B * H): valid when no label can be ignored
};

// workspace layout (floats unless stated): [0] uint32 ticket | [16..16+H) valid counts | [64 ...) block partials
constexpr int kWsCounts = 16;
constexpr int kWsGs = 32;        // per-head gradient scales written by heads_count_kernel (ignore_index form)
constexpr int kWsPartials = 64;

__global__ void heads_count_kernel(const int64_t* __restrict__ labels, int64_t B, int H, int64_t ignore_index,
                                   float* __restrict__ ws, HeadMeta meta, float inv_T, float grad_scale) {
  // one block; exact integer counts of labels != ignore_index per head
  __shared__ int cnt[kMaxHeads];
  if (threadIdx.x < kMaxHeads) cnt[threadIdx.x] = 0;
  __syncthreads();
  int local[kMaxHeads];
#pragma unroll
  for (int h = 0; h < kMaxHeads; ++h) local[h] = 0;
  for (int64_t i = threadIdx.x; i < B * H; i += blockDim.x) {
    const int h = (int)(i % H);
    const bool v = __ldg(labels + i) != ignore_index;
#pragma unroll
    for (int k = 0; k < kMaxHeads; ++k) local[k] += (k == h && v) ? 1 : 0;
  }
#pragma unroll
  for (int h = 0; h < kMaxHeads; ++h) {
    int v = local[h];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0 && h < H && v) atomicAdd(&cnt[h], v);
  }
  __syncthreads();
  if (threadIdx.x < H) {
    const float c = (float)cnt[threadIdx.x];
    ws[kWsCounts + threadIdx.x] = c;
    ws[kWsGs + threadIdx.x] = c > 0.f ? grad_scale * meta.weight[threadIdx.x] * inv_T / (c * (float)H) : 0.f;
  }
}

// kFixed == true: the SM3 head layout (5,3,2,3,3,3,3,2 -> 24 logits, 8 heads) is a compile-time constant, so
// every index computation in the staging loops is a shift/multiply and the per-row softmaxes unroll into
// registers.  kFixed == false: arbitrary layout from `meta` (same algorithm, runtime bounds).
template <typename T, bool kFixed>
__global__ void __launch_bounds__(kHeadThreads)
multihead_ce_kernel(const T* __restrict__ logits, const int64_t* __restrict__ labels, int64_t B, HeadMeta meta,
                    float inv_T, int use_ignore, int64_t ignore_index, float* __restrict__ loss_out,
                    T* __restrict__ dlogits, float grad_scale, float* __restrict__ ws) {
  extern __shared__ float smem[];
  const int C = kFixed ? 24 : meta.C;
  const int H = kFixed ? 8 : meta.H;
  const int ldx = C + 1;                      // +1 float: conflict-free row-per-thread access
  float* xs = smem;                           // [kHeadThreads][C+1]
  int* ys = reinterpret_cast<int*>(smem + kHeadThreads * ldx);  // [kHeadThreads][H+1]
  float* red = reinterpret_cast<float*>(ys + kHeadThreads * (H + 1));  // [8 warps][H]
  __shared__ bool is_last;
  constexpr int kOff[9] = {0, 5, 8, 10, 13, 16, 19, 22, 24};

  const int64_t row0 = (int64_t)blockIdx.x * kHeadThreads;
  const int rows_here = (int)min((int64_t)kHeadThreads, B - row0);
  const int tid = threadIdx.x;

  // ---- coalesced stage-in ----
  const int n_el = rows_here * C;
  const T* src = logits + row0 * C;
  constexpr int V = VecIO<T>::N;
  const bool vec_ok = (((uintptr_t)src & 15u) == 0);
  const int n_vec = vec_ok ? n_el / V : 0;
  for (int v = tid; v < n_vec; v += kHeadThreads) {
    float t[V];
    VecIO<T>::load(src + (size_t)v * V, t);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int e = v * V + i;
      const int r = e / C;
      xs[r * ldx + (e - r * C)] = t[i];
    }
  }
  for (int e = n_vec * V + tid; e < n_el; e += kHeadThreads) { const int r = e / C; xs[r * ldx + (e - r * C)] = to_f32(src[e]); }
  const int64_t* lsrc = labels + row0 * H;
  for (int e = tid; e < rows_here * H; e += kHeadThreads) {
    const int64_t y = __ldg(lsrc + e);
    const int r = e / H, h = e - r * H;
    const int nc = kFixed ? (kOff[h + 1] - kOff[h]) : (meta.offset[h + 1] - meta.offset[h]);
    int yi;
    if (use_ignore && y == ignore_index) yi = -1;
    else yi = (y >= 0 && y < nc) ? (int)y : -2;   // -2: out of range -> NaN loss (torch would device-assert)
    ys[r * (H + 1) + h] = yi;
  }
  __syncthreads();

  // ---- one row per thread ----
  float head_loss[kMaxHeads];
#pragma unroll
  for (int h = 0; h < kMaxHeads; ++h) head_loss[h] = 0.f;
  if (tid < rows_here) {
    float* x = xs + tid * ldx;
    const int* y = ys + tid * (H + 1);
    if constexpr (kFixed) {
#pragma unroll
      for (int h = 0; h < 8; ++h) {
        constexpr int kMaxNc = 5;
        const int o = kOff[h], nc = kOff[h + 1] - kOff[h];
        float v[kMaxNc];
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < kMaxNc; ++c) if (c < nc) { v[c] = x[o + c] * inv_T; mx = fmaxf(mx, v[c]); }
        float se = 0.f;
#pragma unroll
        for (int c = 0; c < kMaxNc; ++c) if (c < nc) { v[c] = __expf(v[c] - mx); se += v[c]; }
        const int yy = y[h];
        const bool valid = yy >= 0;
        const float cnt = use_ignore ? ws[kWsCounts + h] : (float)B;
        const float gs = (valid && cnt > 0.f) ? grad_scale * meta.weight[h] * inv_T / (cnt * 8.0f) : 0.f;
        const float inv_se = __frcp_rn(se);
        float xy = 0.f;
#pragma unroll
        for (int c = 0; c < kMaxNc; ++c) if (c < nc) {
          if (c == yy) xy = x[o + c] * inv_T;
          x[o + c] = (v[c] * inv_se - (c == yy ? 1.f : 0.f)) * gs;
        }
        if (valid) head_loss[h] = mx + __logf(se) - xy;
        if (yy == -2) head_loss[h] = NAN;
      }
    } else {
#pragma unroll 1
      for (int h = 0; h < H; ++h) {
        const int o = meta.offset[h], nc = meta.offset[h + 1] - o;
        float mx = -INFINITY;
        for (int c = 0; c < nc; ++c) mx = fmaxf(mx, x[o + c] * inv_T);
        float se = 0.f;
        for (int c = 0; c < nc; ++c) se += __expf(x[o + c] * inv_T - mx);
        const float lse = mx + __logf(se);
        const int yy = y[h];
        const bool valid = yy >= 0;
        const float cnt = use_ignore ? ws[kWsCounts + h] : (float)B;
        const float gs = (valid && cnt > 0.f) ? grad_scale * meta.weight[h] * inv_T / (cnt * (float)H) : 0.f;
        float hl = 0.f;
        if (valid) hl = lse - x[o + yy] * inv_T;
        if (yy == -2) hl = NAN;
        const float inv_se = 1.0f / se;
        for (int c = 0; c < nc; ++c) {
          const float p = __expf(x[o + c] * inv_T - mx) * inv_se;
          x[o + c] = (p - (c == yy ? 1.f : 0.f)) * gs;
        }
#pragma unroll
        for (int k = 0; k < kMaxHeads; ++k) if (k == h) head_loss[k] = hl;
      }
    }
  }

  // ---- block reduction of the per-head loss sums ----
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int h = 0; h < kMaxHeads; ++h) {
    if (h < H) {
      const float v = warp_sum(head_loss[h]);
      if (lane == 0) red[warp * H + h] = v;
    }
  }
  __syncthreads();
  if (tid < H) {
    float s = 0.f;
    for (int w = 0; w < kHeadThreads / 32; ++w) s += red[w * H + tid];
    ws[kWsPartials + (int64_t)blockIdx.x * H + tid] = s;
  }

  // ---- coalesced stage-out of the gradient ----
  if (dlogits != nullptr) {
    T* dst = dlogits + row0 * C;
    const bool vo = (((uintptr_t)dst & 15u) == 0);
    const int nv = vo ? n_el / V : 0;
    for (int v = tid; v < nv; v += kHeadThreads) {
      float t[V];
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const int e = v * V + i;
        const int r = e / C;
        t[i] = xs[r * ldx + (e - r * C)];
      }
      VecIO<T>::store(dst + (size_t)v * V, t);
    }
    for (int e = nv * V + tid; e < n_el; e += kHeadThreads) { const int r = e / C; dst[e] = from_f32<T>(xs[r * ldx + (e - r * C)]); }
  }

  // ---- last block folds the partials in a fixed order (deterministic) ----
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned t = atomicAdd(reinterpret_cast<unsigned*>(ws), 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    // 16 groups x 16 head lanes: group g sums blocks g, g+16, ... for head (tid & 15); groups are then folded in a
    // fixed order => deterministic for a fixed grid, and parallel enough for 10^4 blocks.
    __shared__ float fold[16][17];
    const int hh = tid & 15, grp = tid >> 4;
    float s = 0.f;
    if (hh < H)
      for (unsigned b = grp; b < gridDim.x; b += 16) s += __ldcg(ws + kWsPartials + (int64_t)b * H + hh);
    fold[grp][hh] = s;
    __syncthreads();
    if (tid < 32) {
      float total = 0.f;
      if (tid < H) {
        float t = 0.f;
        for (int g2 = 0; g2 < 16; ++g2) t += fold[g2][tid];
        const float cnt = use_ignore ? ws[kWsCounts + tid] : (float)B;
        total = meta.weight[tid] * (t / cnt) / (float)H;   // cnt == 0 -> NaN, as torch's mean over nothing
      }
      total = warp_sum(total);    // H <= 16 < 32: every head lives in warp 0
      if (tid == 0) { *loss_out = total; *reinterpret_cast<unsigned*>(ws) = 0u; }
    }
  }
}

// SM3 layout fast path (24 logits = 5,3,2,3,3,3,3,2; 8 int64 labels).  Each warp moves its 32 rows as ONE contiguous
// span (48 or 96 B of logits + 64 B of labels per row) with 16-byte transactions into a per-warp shared-memory slab,
// then every lane works on its own row in registers and the gradient goes back out through the same slab.  No block
// barriers on the data path.
// kAsync: the slab is double buffered and filled with cp.async, two tiles ahead of the softmaxes: the copies hold no
// registers, so each warp keeps ~7 KB in flight through its (long, instruction-bound) compute phase.
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename T> struct CeSlab {
  static constexpr int kRowBytes = 24 * (int)sizeof(T);          // 48 or 96
  static constexpr int kX = 32 * kRowBytes;                      // logits / gradient bytes per warp and buffer
  static constexpr int kY = 32 * 8 * 8;                          // raw int64 labels per warp and buffer
  static constexpr int bytes(bool async) { return (kHeadThreads / 32) * (async ? 2 : 1) * (kX + kY); }
};

template <typename T, bool kAsync>
__global__ void __launch_bounds__(kHeadThreads)
multihead_ce_sm3_kernel(const T* __restrict__ logits, const int64_t* __restrict__ labels, int64_t B, HeadMeta meta,
                        float inv_T, int use_ignore, int64_t ignore_index, float* __restrict__ loss_out,
                        T* __restrict__ dlogits, float grad_scale, float* __restrict__ ws) {
  constexpr int C = 24, H = 8;
  constexpr int kOff[9] = {0, 5, 8, 10, 13, 16, 19, 22, 24};
  constexpr int V = VecIO<T>::N;            // 4 (fp32) or 8 (16-bit)
  constexpr int NV = C / V;                 // 6 or 3 vectors per row
  constexpr int kRowBytes = CeSlab<T>::kRowBytes;
  constexpr int kVX = kRowBytes / 16;       // 16-byte vectors per lane and tile (logits); 4 more for the labels
  constexpr int kBuf = CeSlab<T>::kX + CeSlab<T>::kY;
  extern __shared__ __align__(16) unsigned char slab[];
  __shared__ float red[(kHeadThreads / 32) * H];
  __shared__ bool is_last;
  (void)grad_scale;
  const int tid = threadIdx.x;
  const int lane_ = tid & 31, warp_ = tid >> 5;
  unsigned char* wslab = slab + warp_ * (kAsync ? 2 : 1) * kBuf;

  // per-head gradient scales: host-computed when nothing can be ignored, else from the counting kernel (no barrier)
  float gsc[H];
#pragma unroll
  for (int h = 0; h < H; ++h) gsc[h] = use_ignore ? __ldg(ws + kWsGs + h) : meta.gs[h];
  const float k2 = inv_T * 1.4426950408889634f;     // logits * inv_T, in log2 units

  float head_loss[H];
#pragma unroll
  for (int h = 0; h < H; ++h) head_loss[h] = 0.f;

  const int64_t n_tiles = (B + kHeadThreads - 1) / kHeadThreads;
  // global -> slab `buf` of this warp: the warp's rows of `tile` (warp-coalesced 16-byte transactions)
  auto fetch = [&](int64_t tile, int buf) {
    const int64_t wrow0 = tile * kHeadThreads + warp_ * 32;
    const int wrows = (int)max((int64_t)0, min((int64_t)32, B - wrow0));
    uint4* dx_ = reinterpret_cast<uint4*>(wslab + buf * kBuf);
    uint4* dy_ = reinterpret_cast<uint4*>(wslab + buf * kBuf + CeSlab<T>::kX);
    const uint4* gx = reinterpret_cast<const uint4*>(logits + wrow0 * C);
    const uint4* gy = reinterpret_cast<const uint4*>(labels + wrow0 * H);
    const int nvx = wrows * kVX, nvy = wrows * 4;
    if constexpr (kAsync) {
#pragma unroll
      for (int v = 0; v < kVX; ++v) { const int i = lane_ + 32 * v; if (i < nvx) cp_async16(dx_ + i, gx + i); }
#pragma unroll
      for (int v = 0; v < 4; ++v) { const int i = lane_ + 32 * v; if (i < nvy) cp_async16(dy_ + i, gy + i); }
    } else {
      uint4 rx[kVX], ry[4];
#pragma unroll
      for (int v = 0; v < kVX; ++v) { const int i = lane_ + 32 * v; rx[v] = i < nvx ? __ldg(gx + i) : make_uint4(0u, 0u, 0u, 0u); }
#pragma unroll
      for (int v = 0; v < 4; ++v) { const int i = lane_ + 32 * v; ry[v] = i < nvy ? __ldg(gy + i) : make_uint4(0u, 0u, 0u, 0u); }
#pragma unroll
      for (int v = 0; v < kVX; ++v) dx_[lane_ + 32 * v] = rx[v];
#pragma unroll
      for (int v = 0; v < 4; ++v) dy_[lane_ + 32 * v] = ry[v];
    }
  };
  if constexpr (kAsync) {       // two tiles in flight before the first softmax (empty groups keep the count uniform)
    if ((int64_t)blockIdx.x < n_tiles) fetch(blockIdx.x, 0);
    cp_async_commit();
    if ((int64_t)blockIdx.x + gridDim.x < n_tiles) fetch((int64_t)blockIdx.x + gridDim.x, 1);
    cp_async_commit();
  }

  // persistent loop over 256-row tiles: one loss reduction per CTA instead of one per tile
  int buf = 0;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t wrow0 = tile * kHeadThreads + warp_ * 32;
    const int64_t row = wrow0 + lane_;
    const int wrows = (int)max((int64_t)0, min((int64_t)32, B - wrow0));
    if constexpr (kAsync) cp_async_wait<1>(); else fetch(tile, 0);
    __syncwarp();
    unsigned char* sx = wslab + buf * kBuf;
    const long long* sy = reinterpret_cast<const long long*>(sx + CeSlab<T>::kX);

    if (row < B) {
      float x[C];
      int y[H];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        float t[V];
        const uint4 raw = reinterpret_cast<const uint4*>(sx + lane_ * kRowBytes)[v];
        if constexpr (sizeof(T) == 4) {
          t[0] = __uint_as_float(raw.x); t[1] = __uint_as_float(raw.y); t[2] = __uint_as_float(raw.z); t[3] = __uint_as_float(raw.w);
        } else {
          const unsigned w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if constexpr (std::is_same<T, __nv_bfloat16>::value) {
              t[2 * q] = __uint_as_float(w[q] << 16); t[2 * q + 1] = __uint_as_float(w[q] & 0xFFFF0000u);
            } else {
              const __half2 hh = *reinterpret_cast<const __half2*>(&w[q]);
              const float2 f = __half22float2(hh);
              t[2 * q] = f.x; t[2 * q + 1] = f.y;
            }
          }
        }
#pragma unroll
        for (int i = 0; i < V; ++i) x[v * V + i] = t[i];
      }
      {   // -1: ignored, -2: out of range, else the class id
        const longlong2* yr = reinterpret_cast<const longlong2*>(sy + lane_ * H);
#pragma unroll
        for (int q = 0; q < H / 2; ++q) {
          const longlong2 t = yr[q];
          y[2 * q] = (use_ignore && t.x == ignore_index) ? -1
                     : ((unsigned long long)t.x < (unsigned long long)(kOff[2 * q + 1] - kOff[2 * q]) ? (int)t.x : -2);
          y[2 * q + 1] = (use_ignore && t.y == ignore_index) ? -1
                         : ((unsigned long long)t.y < (unsigned long long)(kOff[2 * q + 2] - kOff[2 * q + 1]) ? (int)t.y : -2);
        }
      }
#pragma unroll
      for (int h = 0; h < H; ++h) {
        constexpr int kMaxNc = 5;
        const int o = kOff[h], nc = kOff[h + 1] - kOff[h];
        float e[kMaxNc];
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < kMaxNc; ++c) if (c < nc) { e[c] = x[o + c] * k2; mx = fmaxf(mx, e[c]); }
        const int yy = y[h];
        const bool in_range = yy >= 0;
        const bool ignored = yy == -1;
        float vy = 0.f, se = 0.f;
#pragma unroll
        for (int c = 0; c < kMaxNc; ++c) if (c < nc) {
          vy = (c == yy) ? e[c] : vy;
          e[c] = ex2f_approx(e[c] - mx);
          se += e[c];
        }
        const float g = in_range ? gsc[h] : 0.f;
        const float rg = rcpf_approx(se) * g;
#pragma unroll
        for (int c = 0; c < kMaxNc; ++c) if (c < nc) x[o + c] = fmaf(e[c], rg, (c == yy) ? -g : 0.f);
        const float hl = 0.6931471805599453f * (mx + lg2f_approx(se) - vy);
        head_loss[h] += in_range ? hl : (ignored ? 0.f : NAN);     // out-of-range label: torch would device-assert
      }
      if (dlogits != nullptr) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          float t[V];
#pragma unroll
          for (int i = 0; i < V; ++i) t[i] = x[v * V + i];
          VecIO<T>::store(reinterpret_cast<T*>(sx + lane_ * kRowBytes) + v * V, t);    // row -> slab
        }
      }
    }
    if (dlogits != nullptr) {
      __syncwarp();
      uint4* gdst = reinterpret_cast<uint4*>(dlogits + wrow0 * C);
      const int nvec = wrows * kRowBytes / 16;
#pragma unroll
      for (int v = 0; v < kRowBytes / 16; ++v) {
        const int i = lane_ + 32 * v;
        if (i < nvec) gdst[i] = reinterpret_cast<const uint4*>(sx)[i];
      }
    }
    __syncwarp();     // slab is reused by a later tile
    if constexpr (kAsync) {
      const int64_t t2 = tile + 2 * (int64_t)gridDim.x;
      if (t2 < n_tiles) fetch(t2, buf);
      cp_async_commit();
      buf ^= 1;
    }
  }
  if constexpr (kAsync) cp_async_wait<0>();

  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int h = 0; h < H; ++h) {
    const float v = warp_sum(head_loss[h]);
    if (lane == 0) red[warp * H + h] = v;
  }
  __syncthreads();
  if (tid < H) {
    float s2 = 0.f;
    for (int w = 0; w < kHeadThreads / 32; ++w) s2 += red[w * H + tid];
    ws[kWsPartials + (int64_t)blockIdx.x * H + tid] = s2;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned t = atomicAdd(reinterpret_cast<unsigned*>(ws), 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    __shared__ float fold[16][17];
    const int hh = tid & 15, grp = tid >> 4;
    float s2 = 0.f;
    if (hh < H)
      for (unsigned b = grp; b < gridDim.x; b += 16) s2 += __ldcg(ws + kWsPartials + (int64_t)b * H + hh);
    fold[grp][hh] = s2;
    __syncthreads();
    if (tid < 32) {
      float total = 0.f;
      if (tid < H) {
        float t = 0.f;
        for (int g2 = 0; g2 < 16; ++g2) t += fold[g2][tid];
        const float cnt = use_ignore ? ws[kWsCounts + tid] : (float)B;
        total = meta.weight[tid] * (t / cnt) / (float)H;
      }
      total = warp_sum(total);
      if (tid == 0) { *loss_out = total; *reinterpret_cast<unsigned*>(ws) = 0u; }
    }
  }
}

// ---------------- K5: BCE with logits ----------------
constexpr int kBceThreads = 256;

// kProd (unweighted form only): the kernel is co-limited by the MUFU pipe (3 transcendentals per element against
// 6 bytes of traffic), so the per-element log is replaced by ONE log per 8 elements of the running product of
// (1 + e^-|x|) -- each factor lies in (1, 2], the product of 8 in (1, 256] -- which leaves 2.125 MUFU ops per element.
template <typename TX, typename TT, bool kHasPW, bool kProd>
__global__ void __launch_bounds__(kBceThreads)
bce_kernel(const TX* __restrict__ x, const TT* __restrict__ t, const float* __restrict__ pos_weight, int64_t n,
           int C, float inv_n, float* __restrict__ loss_out, TX* __restrict__ dx, float grad_scale,
           float* __restrict__ ws, int vec_ok) {
  static_assert(!(kHasPW && kProd), "the product form needs unweighted log terms");
  __shared__ float red[32];
  __shared__ bool is_last;
  float acc = 0.f;
  float prod = 1.f;
  const float gscale = inv_n * grad_scale;
  auto elem = [&](float xv, float tv, int64_t idx) -> float {
    // e = exp(-|x|), r = 1/(1+e):  softplus(-x) = max(-x,0) - log(r),  sigmoid(x) = x >= 0 ? r : 1 - r
    const float e = ex2f_approx(-1.4426950408889634f * fabsf(xv));
    const float w = 1.0f + e;
    const float r = rcpf_approx(w);
    const float sig = xv >= 0.f ? r : 1.0f - r;
    if constexpr (kProd) {
      prod *= w;
      acc += fmaf(1.f - tv, xv, fmaxf(-xv, 0.f));
      return (sig - tv) * gscale;
    } else {
      const float sp = fmaxf(-xv, 0.f) - 0.6931471805599453f * lg2f_approx(r);
      if constexpr (kHasPW) {
        const float lw = fmaf(__ldg(pos_weight + (idx % C)) - 1.f, tv, 1.f);
        acc += fmaf(1.f - tv, xv, lw * sp);
        return ((1.f - tv) - lw * (1.f - sig)) * gscale;
      } else {
        acc += fmaf(1.f - tv, xv, sp);
        return (sig - tv) * gscale;
      }
    }
  };
  auto fold = [&]() {       // product form: one log per 8 elements
    if constexpr (kProd) { acc = fmaf(0.6931471805599453f, lg2f_approx(prod), acc); prod = 1.f; }
  };
  const int64_t stride = (int64_t)gridDim.x * kBceThreads;
  const int64_t gid = (int64_t)blockIdx.x * kBceThreads + threadIdx.x;
  const int64_t n8 = vec_ok ? n / 8 : 0;
  auto load8x = [&](int64_t v, float (&xv)[8]) {
    if constexpr (sizeof(TX) == 4) {
      float a[4], b[4];
      VecIO<float>::load((const float*)x + v * 8, a); VecIO<float>::load((const float*)x + v * 8 + 4, b);
#pragma unroll
      for (int i = 0; i < 4; ++i) { xv[i] = a[i]; xv[4 + i] = b[i]; }
    } else { VecIO<TX>::load(x + v * 8, xv); }
  };
  auto load8t = [&](int64_t v, float (&tv)[8]) {
    if constexpr (sizeof(TT) == 4) {
      float a[4], b[4];
      VecIO<float>::load((const float*)t + v * 8, a); VecIO<float>::load((const float*)t + v * 8 + 4, b);
#pragma unroll
      for (int i = 0; i < 4; ++i) { tv[i] = a[i]; tv[4 + i] = b[i]; }
    } else { VecIO<TT>::load(t + v * 8, tv); }
  };
  auto store8 = [&](int64_t v, const float (&g)[8]) {
    if (dx == nullptr) return;
    if constexpr (sizeof(TX) == 4) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = g[i]; b[i] = g[4 + i]; }
      VecIO<float>::store((float*)dx + v * 8, a); VecIO<float>::store((float*)dx + v * 8 + 4, b);
    } else { VecIO<TX>::store(dx + v * 8, g); }
  };
  for (int64_t v = gid; v < n8; v += 2 * stride) {       // two independent chunks in flight per thread
    const int64_t v2 = v + stride;
    const bool two = v2 < n8;
    float xa[8], ta[8], xb[8], tb[8], g[8];
    load8x(v, xa); load8t(v, ta);
    if (two) { load8x(v2, xb); load8t(v2, tb); }
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] = elem(xa[i], ta[i], v * 8 + i);
    fold();
    store8(v, g);
    if (two) {
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] = elem(xb[i], tb[i], v2 * 8 + i);
      fold();
      store8(v2, g);
    }
  }
  for (int64_t e = n8 * 8 + gid; e < n; e += stride) {
    const float g = elem(to_f32(x[e]), to_f32(t[e]), e);
    fold();
    if (dx != nullptr) dx[e] = from_f32<TX>(g);
  }
  const float bs = block_sum(acc, red);
  if (threadIdx.x == 0) ws[kWsPartials + blockIdx.x] = bs;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned tk = atomicAdd(reinterpret_cast<unsigned*>(ws), 1u);
    is_last = (tk == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    float s = 0.f;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += kBceThreads) s += __ldcg(ws + kWsPartials + b);
    s = block_sum(s, red);   // fixed tree for a fixed grid => deterministic
    if (threadIdx.x == 0) { *loss_out = s * inv_n; *reinterpret_cast<unsigned*>(ws) = 0u; }
  }
}

// Experimental K5 variant (SM3_BCE_VARIANT=2|3, 16-bit logits and targets, no pos_weight, n % 8 == 0): FOUR 8-element
// chunks per thread are requested (packed, 8 registers per chunk) before the first one is consumed -- twice the bytes in
// flight of bce_kernel at the same occupancy; kNewton additionally takes 1 / (1 + e) from a linear seed + 2 Newton steps
// on the FMA pipe (1.2e-5 relative, below 16-bit output rounding) so that ONE MUFU op per element is left (+ one log
// per 8).  Not a default until measured.
template <typename T>
__device__ __forceinline__ void unpack8_16(const uint4& r, float (&o)[8]) {
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if constexpr (std::is_same<T, __nv_bfloat16>::value) {
      o[2 * q] = __uint_as_float(w[q] << 16); o[2 * q + 1] = __uint_as_float(w[q] & 0xFFFF0000u);
    } else {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[q]));
      o[2 * q] = f.x; o[2 * q + 1] = f.y;
    }
  }
}

template <typename TX, typename TT, bool kNewton>
__global__ void __launch_bounds__(kBceThreads)
bce_deep_kernel(const TX* __restrict__ x, const TT* __restrict__ t, int64_t n8, float inv_n, float* __restrict__ loss_out,
                TX* __restrict__ dx, float grad_scale, float* __restrict__ ws) {
  static_assert(sizeof(TX) == 2 && sizeof(TT) == 2, "16-bit inputs only");
  constexpr int U = 4;
  __shared__ float red[32];
  __shared__ bool is_last;
  float acc = 0.f;
  const float gscale = inv_n * grad_scale;
  const int64_t stride = (int64_t)gridDim.x * kBceThreads;
  const int64_t gid = (int64_t)blockIdx.x * kBceThreads + threadIdx.x;
  const uint4* xv4 = reinterpret_cast<const uint4*>(x);
  const uint4* tv4 = reinterpret_cast<const uint4*>(t);
  for (int64_t v0 = gid; v0 < n8; v0 += U * stride) {
    uint4 rx[U], rt[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = v0 + u * stride;
      if (v < n8) { rx[u] = __ldg(xv4 + v); rt[u] = __ldg(tv4 + v); }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t v = v0 + u * stride;
      if (v < n8) {
        float xs[8], ts[8], g[8];
        unpack8_16<TX>(rx[u], xs);
        unpack8_16<TT>(rt[u], ts);
        float prod = 1.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float e = ex2f_approx(-1.4426950408889634f * fabsf(xs[i]));
          const float w = 1.0f + e;
          float r;
          if constexpr (kNewton) {
            r = fmaf(w, -0.47058823529f, 1.41176470588f);      // 24/17 - 8/17 w : |rel err| <= 1/17 on [1, 2]
            r = r * fmaf(-w, r, 2.0f);
            r = r * fmaf(-w, r, 2.0f);
          } else {
            r = rcpf_approx(w);
          }
          const float sig = xs[i] >= 0.f ? r : 1.0f - r;
          prod *= w;
          acc += fmaf(1.f - ts[i], xs[i], fmaxf(-xs[i], 0.f));
          g[i] = (sig - ts[i]) * gscale;
        }
        acc = fmaf(0.6931471805599453f, lg2f_approx(prod), acc);
        if (dx != nullptr) VecIO<TX>::store(dx + v * 8, g);
      }
    }
  }
  const float bs = block_sum(acc, red);
  if (threadIdx.x == 0) ws[kWsPartials + blockIdx.x] = bs;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned tk = atomicAdd(reinterpret_cast<unsigned*>(ws), 1u);
    is_last = (tk == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    float s = 0.f;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += kBceThreads) s += __ldcg(ws + kWsPartials + b);
    s = block_sum(s, red);
    if (threadIdx.x == 0) { *loss_out = s * inv_n; *reinterpret_cast<unsigned*>(ws) = 0u; }
  }
}

// ---------------- InfoNCE tail kernels ----------------
__global__ void infonce_finalize_kernel(const float* __restrict__ partial, int n_partials, int64_t rows, float inv_T,
                                        float* __restrict__ neg_sum, float* __restrict__ lse_neg,
                                        const float* __restrict__ extra) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  const float s = (extra ? extra[i] : 0.f) + fold_row_partials(partial, n_partials, rows, i);
  neg_sum[i] = s;
  lse_neg[i] = inv_T + logf(s);     // s == 0 (no negatives, N == 1) -> -inf, CE([pos,-inf],0) = 0
}

__device__ __forceinline__ float ce_stat_term(float pos, float lse, float scale, float& g) {
  const float x = lse - pos;
  const float e = __expf(-fabsf(x));
  const float r = __frcp_rn(1.0f + e);
  g = scale * (x >= 0.f ? r : e * r);                      // scale * sigmoid(x)
  return fmaxf(x, 0.f) + log1pf(e);                        // log(e^pos + e^lse) - pos
}

__global__ void __launch_bounds__(1024)
infonce_loss_kernel(const float* __restrict__ pos, const float* __restrict__ lse_neg, int64_t rows, float scale,
                    float* __restrict__ loss, int accumulate, float* __restrict__ g_pos, float* __restrict__ g_lse) {
  __shared__ float red[32];
  float acc = 0.f;
  const bool vec = ((((uintptr_t)pos | (uintptr_t)lse_neg | (uintptr_t)g_pos | (uintptr_t)g_lse) & 15u) == 0);
  const int64_t n4 = vec ? rows / 4 : 0;
  for (int64_t i = threadIdx.x; i < n4; i += 2 * blockDim.x) {
    const int64_t j = i + blockDim.x;
    const bool two = j < n4;
    const float4 p0 = reinterpret_cast<const float4*>(pos)[i];
    const float4 l0 = reinterpret_cast<const float4*>(lse_neg)[i];
    float4 p1 = p0, l1 = l0;
    if (two) { p1 = reinterpret_cast<const float4*>(pos)[j]; l1 = reinterpret_cast<const float4*>(lse_neg)[j]; }
    float4 g0, g1;
    acc += ce_stat_term(p0.x, l0.x, scale, g0.x) + ce_stat_term(p0.y, l0.y, scale, g0.y) +
           ce_stat_term(p0.z, l0.z, scale, g0.z) + ce_stat_term(p0.w, l0.w, scale, g0.w);
    if (g_lse) reinterpret_cast<float4*>(g_lse)[i] = g0;
    if (g_pos) reinterpret_cast<float4*>(g_pos)[i] = make_float4(-g0.x, -g0.y, -g0.z, -g0.w);
    if (two) {
      acc += ce_stat_term(p1.x, l1.x, scale, g1.x) + ce_stat_term(p1.y, l1.y, scale, g1.y) +
             ce_stat_term(p1.z, l1.z, scale, g1.z) + ce_stat_term(p1.w, l1.w, scale, g1.w);
      if (g_lse) reinterpret_cast<float4*>(g_lse)[j] = g1;
      if (g_pos) reinterpret_cast<float4*>(g_pos)[j] = make_float4(-g1.x, -g1.y, -g1.z, -g1.w);
    }
  }
  for (int64_t i = n4 * 4 + threadIdx.x; i < rows; i += blockDim.x) {
    float g;
    acc += ce_stat_term(pos[i], lse_neg[i], scale, g);
    if (g_lse) g_lse[i] = g;
    if (g_pos) g_pos[i] = -g;
  }
  const float s = block_sum(acc, red);
  if (threadIdx.x == 0) *loss = (accumulate ? *loss : 0.f) + s * scale;
}

}  // namespace

int infonce_finalize_launch(const float* partial_sums, int n_partials, int64_t rows, float inv_T, float* neg_sum,
                            float* lse_neg, cudaStream_t st, const float* extra) {
  if (rows == 0) return SM3_OK;
  infonce_finalize_kernel<<<(unsigned)((rows + 255) / 256), 256, 0, st>>>(partial_sums, n_partials, rows, inv_T,
                                                                           neg_sum, lse_neg, extra);
  SM3_CHECK_CUDA(cudaGetLastError());
  return SM3_OK;
}

int infonce_loss_launch(const float* pos, const float* lse_neg, int64_t rows, float scale, float* loss, int accumulate,
                        float* g_pos, float* g_lse, cudaStream_t st) {
  infonce_loss_kernel<<<1, 1024, 0, st>>>(pos, lse_neg, rows, scale, loss, accumulate, g_pos, g_lse);
  SM3_CHECK_CUDA(cudaGetLastError());
  return SM3_OK;
}

}  // namespace sm3

// =====================================================================================================
// C ABI
// =====================================================================================================
using namespace sm3;

extern "C" size_t sm3_multihead_ce_workspace_bytes(int64_t B, int H) {
  const int64_t blocks = (B + kHeadThreads - 1) / kHeadThreads;
  return (size_t)(kWsPartials + blocks * (H > 0 ? H : 1)) * sizeof(float);
}

extern "C" int sm3_multihead_ce(const void* logits, int dtype, const int64_t* labels, int64_t B, int H,
                                const int* class_counts_host, const float* weights_host, float inv_T,
                                int use_ignore_index, int64_t ignore_index, float* loss, void* dlogits,
                                float grad_scale, void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SM3_REQUIRE(logits && labels && loss && workspace && class_counts_host, SM3_ERR_SHAPE, "multihead_ce: null pointer");
  SM3_REQUIRE(dtype_ok(dtype), SM3_ERR_DTYPE, "multihead_ce: bad dtype %d", dtype);
  SM3_REQUIRE(H >= 1 && H <= kMaxHeads, SM3_ERR_SHAPE, "multihead_ce: H=%d not in [1,%d]", H, kMaxHeads);
  SM3_REQUIRE(B >= 1, SM3_ERR_SHAPE, "multihead_ce: empty batch");
  HeadMeta meta;
  meta.H = H;
  meta.offset[0] = 0;
  for (int h = 0; h < H; ++h) {
    SM3_REQUIRE(class_counts_host[h] >= 1, SM3_ERR_SHAPE, "multihead_ce: head %d has %d classes", h, class_counts_host[h]);
    meta.offset[h + 1] = meta.offset[h] + class_counts_host[h];
    meta.weight[h] = weights_host ? weights_host[h] : 1.0f;
    meta.gs[h] = grad_scale * meta.weight[h] * inv_T / ((float)B * (float)H);
  }
  meta.C = meta.offset[H];
  SM3_REQUIRE(meta.C <= kMaxClassesTotal, SM3_ERR_SHAPE, "multihead_ce: %d total classes > %d", meta.C, kMaxClassesTotal);
  SM3_REQUIRE(workspace_bytes >= sm3_multihead_ce_workspace_bytes(B, H), SM3_ERR_WORKSPACE, "multihead_ce: workspace too small");
  float* ws = (float*)workspace;
  SM3_CHECK_CUDA(cudaMemsetAsync(ws, 0, kWsPartials * sizeof(float), st));
  if (use_ignore_index) {
    heads_count_kernel<<<1, 1024, 0, st>>>(labels, B, H, ignore_index, ws, meta, inv_T, grad_scale);
    SM3_CHECK_CUDA(cudaGetLastError());
  }
  const unsigned grid = (unsigned)((B + kHeadThreads - 1) / kHeadThreads);
  const size_t smem = (size_t)kHeadThreads * (meta.C + 1) * 4 + (size_t)kHeadThreads * (H + 1) * 4 + (kHeadThreads / 32) * H * 4;
  static const int kSm3Layout[8] = {5, 3, 2, 3, 3, 3, 3, 2};
  bool fixed = (H == 8);
  for (int h = 0; fixed && h < 8; ++h) fixed = (class_counts_host[h] == kSm3Layout[h]);
  SM3_DISPATCH_DTYPE(dtype, T, {
    if (fixed && aligned16(logits) && aligned16(labels) && (dlogits == nullptr || aligned16(dlogits))) {
      const unsigned pgrid = grid < (unsigned)num_sms() * 8u ? grid : (unsigned)num_sms() * 8u;
      // SM3_CE_VARIANT: 0 = load -> compute -> store per tile (default), 1 = cp.async ring.  Measured on B200 at
      // B = 4M bf16 rows: 115 us (0.89 of the copy rate) vs 141 us -- the kernel is instruction-issue bound (~500
      // instructions per row), and the ring's extra shared memory costs a resident CTA per SM.
      const char* ev = getenv("SM3_CE_VARIANT");
      if (ev && ev[0] == '1') {
        const int smem3 = CeSlab<T>::bytes(true);
        static const int per_sm = [&] {
          int n = 0;
          cudaFuncSetAttribute(multihead_ce_sm3_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem3);
          if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, multihead_ce_sm3_kernel<T, true>, kHeadThreads, smem3) != cudaSuccess || n < 1) n = 2;
          return n;
        }();
        const unsigned cap = (unsigned)(num_sms() * per_sm);
        multihead_ce_sm3_kernel<T, true><<<grid < cap ? grid : cap, kHeadThreads, smem3, st>>>((const T*)logits, labels, B,
            meta, inv_T, use_ignore_index, ignore_index, loss, (T*)dlogits, grad_scale, ws);
      } else {
        static const int per_sm0 = resident_ctas(multihead_ce_sm3_kernel<T, false>, kHeadThreads, CeSlab<T>::bytes(false));
        const unsigned cap0 = (unsigned)(num_sms() * per_sm0);      // one wave of persistent CTAs
        multihead_ce_sm3_kernel<T, false><<<grid < cap0 ? grid : cap0, kHeadThreads, CeSlab<T>::bytes(false), st>>>(
            (const T*)logits, labels, B, meta, inv_T, use_ignore_index, ignore_index, loss, (T*)dlogits, grad_scale, ws);
      }
    } else if (fixed) {
      SM3_CHECK_CUDA(cudaFuncSetAttribute(multihead_ce_kernel<T, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      multihead_ce_kernel<T, true><<<grid, kHeadThreads, smem, st>>>((const T*)logits, labels, B, meta, inv_T,
          use_ignore_index, ignore_index, loss, (T*)dlogits, grad_scale, ws);
    } else {
      SM3_CHECK_CUDA(cudaFuncSetAttribute(multihead_ce_kernel<T, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      multihead_ce_kernel<T, false><<<grid, kHeadThreads, smem, st>>>((const T*)logits, labels, B, meta, inv_T,
          use_ignore_index, ignore_index, loss, (T*)dlogits, grad_scale, ws);
    }
  });
  SM3_CHECK_CUDA(cudaGetLastError());
  return SM3_OK;
}

static unsigned bce_grid(int64_t n) {
  const int64_t want = (n / 16 + kBceThreads - 1) / kBceThreads + 1;
  const int64_t cap = (int64_t)num_sms() * 8;
  return (unsigned)(want < cap ? (want < 1 ? 1 : want) : cap);
}

extern "C" size_t sm3_bce_workspace_bytes(int64_t B, int C) {
  (void)B; (void)C;
  return (size_t)(kWsPartials + 148 * 8 * 2) * sizeof(float);
}

extern "C" int sm3_bce_logits(const void* x, int x_dtype, const void* t, int t_dtype, const float* pos_weight,
                              int64_t B, int C, float* loss, void* dx, float grad_scale, void* workspace,
                              size_t workspace_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SM3_REQUIRE(x && t && loss && workspace, SM3_ERR_SHAPE, "bce: null pointer");
  SM3_REQUIRE(dtype_ok(x_dtype) && dtype_ok(t_dtype), SM3_ERR_DTYPE, "bce: bad dtype");
  SM3_REQUIRE(B >= 1 && C >= 1, SM3_ERR_SHAPE, "bce: empty input");
  const int64_t n = B * (int64_t)C;
  unsigned grid = bce_grid(n);
  SM3_REQUIRE(workspace_bytes >= (size_t)(kWsPartials + grid) * sizeof(float), SM3_ERR_WORKSPACE, "bce: workspace too small");
  float* ws = (float*)workspace;
  SM3_CHECK_CUDA(cudaMemsetAsync(ws, 0, 16, st));
  const int vec_ok = aligned16(x) && aligned16(t) && (dx == nullptr || aligned16(dx));
  const float inv_n = 1.0f / (float)n;
  // SM3_BCE_VARIANT: 0 = one log per element, 1 = one log per 8 elements (default; B200, 4M x 24 bf16: 125 -> 116 us)
  //                  2 / 3 = experimental deep-prefetch kernels (bce_deep_kernel), opt-in only
  const char* ev = getenv("SM3_BCE_VARIANT");
  const bool prod = !(ev && ev[0] == '0');
  const int deep = (ev && (ev[0] == '2' || ev[0] == '3')) ? ev[0] - '0' : 0;
  if (deep && pos_weight == nullptr && vec_ok && n % 8 == 0 && x_dtype != SM3_F32 && t_dtype != SM3_F32) {
    const int64_t n8 = n / 8;
#define SM3_BCE_DEEP(TX, TT)                                                                                         \
    do {                                                                                                             \
      const int per_sm = deep == 3 ? resident_ctas(bce_deep_kernel<TX, TT, true>, kBceThreads)                       \
                                   : resident_ctas(bce_deep_kernel<TX, TT, false>, kBceThreads);                     \
      int64_t g = (n8 + kBceThreads * 4 - 1) / (kBceThreads * 4);                                                    \
      const int64_t cap = (int64_t)num_sms() * per_sm;                                                               \
      if (g > cap) g = cap;                                                                                          \
      if (g > (int64_t)grid) g = grid;          /* the workspace was sized for `grid` CTAs */                        \
      if (deep == 3) bce_deep_kernel<TX, TT, true><<<(unsigned)g, kBceThreads, 0, st>>>((const TX*)x, (const TT*)t, n8, inv_n, loss, (TX*)dx, grad_scale, ws); \
      else bce_deep_kernel<TX, TT, false><<<(unsigned)g, kBceThreads, 0, st>>>((const TX*)x, (const TT*)t, n8, inv_n, loss, (TX*)dx, grad_scale, ws); \
    } while (0)
    if (x_dtype == SM3_BF16 && t_dtype == SM3_BF16) SM3_BCE_DEEP(__nv_bfloat16, __nv_bfloat16);
    else if (x_dtype == SM3_BF16) SM3_BCE_DEEP(__nv_bfloat16, __half);
    else if (t_dtype == SM3_BF16) SM3_BCE_DEEP(__half, __nv_bfloat16);
    else SM3_BCE_DEEP(__half, __half);
#undef SM3_BCE_DEEP
    SM3_CHECK_CUDA(cudaGetLastError());
    return SM3_OK;
  }
  SM3_DISPATCH_DTYPE(x_dtype, TX, SM3_DISPATCH_DTYPE(t_dtype, TT, {
    if (pos_weight != nullptr)
      bce_kernel<TX, TT, true, false><<<grid, kBceThreads, 0, st>>>((const TX*)x, (const TT*)t, pos_weight, n, C, inv_n,
                                                                    loss, (TX*)dx, grad_scale, ws, vec_ok);
    else if (prod) {
      static const int per_sm = resident_ctas(bce_kernel<TX, TT, false, true>, kBceThreads);
      const unsigned cap = (unsigned)(num_sms() * per_sm);          // one wave of persistent CTAs
      if (grid > cap) grid = cap;
      bce_kernel<TX, TT, false, true><<<grid, kBceThreads, 0, st>>>((const TX*)x, (const TT*)t, pos_weight, n, C, inv_n,
                                                                    loss, (TX*)dx, grad_scale, ws, vec_ok);
    } else
      bce_kernel<TX, TT, false, false><<<grid, kBceThreads, 0, st>>>((const TX*)x, (const TT*)t, pos_weight, n, C, inv_n,
                                                                     loss, (TX*)dx, grad_scale, ws, vec_ok);
  }));
  SM3_CHECK_CUDA(cudaGetLastError());
  return SM3_OK;
}

extern "C" int sm3_infonce_loss(const float* pos, const float* lse_neg, int64_t rows, float scale, float* loss,
                                int accumulate, float* g_pos, float* g_lse, void* stream) {
  SM3_REQUIRE(pos && lse_neg && loss, SM3_ERR_SHAPE, "infonce_loss: null pointer");
  return infonce_loss_launch(pos, lse_neg, rows, scale, loss, accumulate, g_pos, g_lse, (cudaStream_t)stream);
}
