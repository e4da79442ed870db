// N1: similarity top-k retrieval.  Replaces `sim_matrix = query @ bank.T; sim_matrix.topk(k)` of
// KNNOnlineEvaluator.predict (reference src/models/evaluator.py:61-63) and the in-batch "positive is the top-1
// non-self neighbour" probe sketched at tools/backbone_train.py:103-105.  The [Bq, Nb] similarity matrix is never
// written: each CTA owns 4 query rows, streams the bank in 1024-row chunks (one warp per bank row, coalesced
// 128-bit loads, fp32 accumulate), and keeps a sorted running top-K per query in shared memory
// (bitonic sort of the chunk + bitonic merge).  Ties resolve to the lower bank index (64-bit keys).
#include "common.cuh"

namespace sm3 {
namespace {

constexpr int kQT = 4;          // queries per CTA
constexpr int kCH = 1024;       // bank rows per chunk
constexpr int kTopkThreads = 256;
constexpr int kMaxK = 256;

__device__ __forceinline__ unsigned long long make_key(float v, unsigned idx) {
  unsigned u = __float_as_uint(v);
  u ^= (u >> 31) ? 0xFFFFFFFFu : 0x80000000u;          // order-preserving map float -> uint
  return ((unsigned long long)u << 32) | (unsigned long long)(~idx);   // equal values: lower idx = larger key
}
__device__ __forceinline__ float key_value(unsigned long long k) {
  unsigned u = (unsigned)(k >> 32);
  u ^= (u >> 31) ? 0x80000000u : 0xFFFFFFFFu;
  return __uint_as_float(u);
}

// in-place bitonic sort (descending) of n = power of two keys in shared memory, whole CTA
__device__ void bitonic_sort_desc(unsigned long long* a, int n) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (int t = threadIdx.x; t < n / 2; t += kTopkThreads) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const unsigned long long x = a[lo], y = a[hi];
        if ((x < y) == desc) { a[lo] = y; a[hi] = x; }
      }
    }
  }
  __syncthreads();
}
// a[0..n) is bitonic -> sorted descending
__device__ void bitonic_merge_desc(unsigned long long* a, int n) {
  for (int stride = n >> 1; stride > 0; stride >>= 1) {
    __syncthreads();
    for (int t = threadIdx.x; t < n / 2; t += kTopkThreads) {
      const int lo = 2 * t - (t & (stride - 1));
      const int hi = lo + stride;
      const unsigned long long x = a[lo], y = a[hi];
      if (x < y) { a[lo] = y; a[hi] = x; }
    }
  }
  __syncthreads();
}

template <typename T>
__global__ void __launch_bounds__(kTopkThreads)
sim_topk_kernel(const T* __restrict__ query, const T* __restrict__ bank, int64_t n_query, int64_t n_bank, int D, int k,
                int K, int64_t exclude_self_offset, float* __restrict__ vals, int64_t* __restrict__ idx) {
  extern __shared__ unsigned char smem_raw[];
  unsigned long long* run = reinterpret_cast<unsigned long long*>(smem_raw);   // [kQT][K] running top-K (sorted desc)
  unsigned long long* cand = run + kQT * K;                                    // [kCH]
  float* sims = reinterpret_cast<float*>(cand + kCH);                          // [kQT][kCH]
  float* q = sims + kQT * kCH;                                                 // [kQT][D]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t q0 = (int64_t)blockIdx.x * kQT;
  const int nq = (int)min((int64_t)kQT, n_query - q0);

  for (int i = tid; i < kQT * D; i += kTopkThreads) {
    const int qi = i / D, d = i - qi * D;
    q[i] = qi < nq ? to_f32(query[(q0 + qi) * D + d]) : 0.f;
  }
  for (int i = tid; i < kQT * K; i += kTopkThreads) run[i] = 0ull;
  __syncthreads();

  constexpr int V = VecIO<T>::N;
  const bool vec = (D % V == 0) && ((((uintptr_t)bank) & 15u) == 0);
  for (int64_t c0 = 0; c0 < n_bank; c0 += kCH) {
    const int rows = (int)min((int64_t)kCH, n_bank - c0);
    // ---- dot products: one warp per bank row, kQT queries at once ----
    for (int r = warp; r < rows; r += kTopkThreads / 32) {
      const T* brow = bank + (c0 + r) * D;
      float acc[kQT];
#pragma unroll
      for (int qi = 0; qi < kQT; ++qi) acc[qi] = 0.f;
      if (vec) {
        for (int ch = lane; ch < D / V; ch += 32) {
          float b[V];
          VecIO<T>::load(brow + ch * V, b);
#pragma unroll
          for (int qi = 0; qi < kQT; ++qi) {
            const float* qq = q + qi * D + ch * V;
#pragma unroll
            for (int e = 0; e < V; ++e) acc[qi] = fmaf(b[e], qq[e], acc[qi]);
          }
        }
      } else {
        for (int d = lane; d < D; d += 32) {
          const float b = to_f32(brow[d]);
#pragma unroll
          for (int qi = 0; qi < kQT; ++qi) acc[qi] = fmaf(b, q[qi * D + d], acc[qi]);
        }
      }
#pragma unroll
      for (int qi = 0; qi < kQT; ++qi) {
        const float s = warp_sum(acc[qi]);
        if (lane == 0) sims[qi * kCH + r] = s;
      }
    }
    __syncthreads();
    // ---- per query: keys -> sort chunk -> merge into the running list ----
    for (int qi = 0; qi < nq; ++qi) {
      const int64_t self = exclude_self_offset >= 0 ? exclude_self_offset + q0 + qi : -1;
      for (int i = tid; i < kCH; i += kTopkThreads) {
        const int64_t j = c0 + i;
        cand[i] = (i < rows && j != self) ? make_key(sims[qi * kCH + i], (unsigned)j) : 0ull;
      }
      bitonic_sort_desc(cand, kCH);
      unsigned long long* rq = run + qi * K;
      // top-K of (run U cand): elementwise max of run and reversed cand-top-K is bitonic
      for (int i = tid; i < K; i += kTopkThreads) {
        const unsigned long long a = rq[i], b = cand[K - 1 - i];
        rq[i] = a > b ? a : b;
      }
      bitonic_merge_desc(rq, K);
    }
    __syncthreads();
  }
  for (int i = tid; i < nq * k; i += kTopkThreads) {
    const int qi = i / k, r = i - qi * k;
    const unsigned long long key = run[qi * K + r];
    const bool ok = key != 0ull;
    vals[(q0 + qi) * k + r] = ok ? key_value(key) : -INFINITY;
    idx[(q0 + qi) * k + r] = ok ? (int64_t)(~(unsigned)(key & 0xFFFFFFFFull)) : -1;
  }
}


// ---------------------------------------------------------------------------------------------------------------
// Tiled form (sm3_sim_topk_ws): the dot products are a register-tiled FP32 contraction (32 queries x 128 bank rows per
// CTA step, 4 x 4 outputs per thread, operands staged k-major in shared memory) instead of one warp per bank row, and the
// selection is threshold-filtered: a value only enters a query's candidate buffer if its key beats the query's current
// K-th best, and a buffer is sorted + merged into the running list (by one warp) only when it may overflow -- after the
// first few tiles that is rare (expected candidates per query ~ K (1 + ln(rows / K))).  The bank is split across CTAs so
// that few queries still fill the machine; a second kernel merges the per-split lists.  Same 64-bit keys as above, so the
// result is the exact top-k with ties resolved to the lower bank index.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kQ2 = 32, kB2 = 128, kKC = 32, kT2 = 256, kCap = 256;

__device__ void warp_bitonic_sort_desc(unsigned long long* a, int n, int lane) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      __syncwarp();
      for (int t = lane; t < n / 2; t += 32) {
        const int lo = 2 * t - (t & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const unsigned long long x = a[lo], y = a[hi];
        if ((x < y) == desc) { a[lo] = y; a[hi] = x; }
      }
    }
  }
  __syncwarp();
}
__device__ void warp_bitonic_merge_desc(unsigned long long* a, int n, int lane) {
  for (int stride = n >> 1; stride > 0; stride >>= 1) {
    __syncwarp();
    for (int t = lane; t < n / 2; t += 32) {
      const int lo = 2 * t - (t & (stride - 1));
      const int hi = lo + stride;
      const unsigned long long x = a[lo], y = a[hi];
      if (x < y) { a[lo] = y; a[hi] = x; }
    }
  }
  __syncwarp();
}
// run[0..K) (sorted descending) <- top-K of run U cand[0..c); one warp; cand is clobbered
__device__ void warp_merge_candidates(unsigned long long* run, unsigned long long* cand, int c, int K, int lane) {
  int n = 32;
  while (n < c) n <<= 1;
  for (int i = c + lane; i < n; i += 32) cand[i] = 0ull;
  warp_bitonic_sort_desc(cand, n, lane);
  for (int i = lane; i < K; i += 32) {
    const unsigned long long a = run[i], b = (K - 1 - i) < n ? cand[K - 1 - i] : 0ull;
    run[i] = a > b ? a : b;
  }
  warp_bitonic_merge_desc(run, K, lane);
}

template <typename T>
__global__ void __launch_bounds__(kT2)
sim_topk_tile_kernel(const T* __restrict__ query, const T* __restrict__ bank, int64_t n_query, int64_t n_bank, int D,
                     int K, int64_t exclude_self_offset, int64_t rows_per_split, int vec,
                     unsigned long long* __restrict__ lists) {
  extern __shared__ unsigned char smem_raw[];
  float* Qs = reinterpret_cast<float*>(smem_raw);                         // [kKC][kQ2]
  float* Bs = Qs + kKC * kQ2;                                             // [kKC][kB2]
  unsigned long long* run = reinterpret_cast<unsigned long long*>(Bs + kKC * kB2);   // [kQ2][K]
  unsigned long long* cand = run + kQ2 * K;                               // [kQ2][kCap]
  unsigned long long* tau = cand + kQ2 * kCap;                            // [kQ2]
  int* cnt = reinterpret_cast<int*>(tau + kQ2);                           // [kQ2]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = lane, ty = warp;                                         // 4 bank columns tx + 32 j | 4 queries 4 ty + i
  const int64_t q0 = (int64_t)blockIdx.x * kQ2;
  const int64_t b_begin = (int64_t)blockIdx.y * rows_per_split;
  const int64_t b_end = min(n_bank, b_begin + rows_per_split);
  constexpr int V = VecIO<T>::N;

  for (int i = tid; i < kQ2 * K; i += kT2) run[i] = 0ull;
  if (tid < kQ2) { tau[tid] = 0ull; cnt[tid] = 0; }
  __syncthreads();

  // operand tiles travel global -> registers -> shared memory: the loads of the NEXT k-chunk are issued before the FMAs
  // of the current one, so their latency is covered by compute (vector path; the scalar path stages directly)
  constexpr int NQV = (kQ2 * (kKC / V) + kT2 - 1) / kT2, NBV = kB2 * (kKC / V) / kT2;
  float qreg[NQV][V], breg[NBV][V];
  auto g_load = [&](int64_t b0, int k0) {
#pragma unroll
    for (int i = 0; i < NQV; ++i) {
      const int v = tid + i * kT2;
      const int r = v % kQ2, kv = v / kQ2;
      if (v < kQ2 * (kKC / V) && q0 + r < n_query && k0 + kv * V < D) VecIO<T>::load(query + (q0 + r) * D + k0 + kv * V, qreg[i]);
      else {
#pragma unroll
        for (int e = 0; e < V; ++e) qreg[i][e] = 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < NBV; ++i) {
      const int v = tid + i * kT2;
      const int r = v % kB2, kv = v / kB2;
      if (b0 + r < b_end && k0 + kv * V < D) VecIO<T>::load(bank + (b0 + r) * D + k0 + kv * V, breg[i]);
      else {
#pragma unroll
        for (int e = 0; e < V; ++e) breg[i][e] = 0.f;
      }
    }
  };
  auto s_store = [&]() {
#pragma unroll
    for (int i = 0; i < NQV; ++i) {
      const int v = tid + i * kT2;
      const int r = v % kQ2, kv = v / kQ2;
      if (v < kQ2 * (kKC / V)) {
#pragma unroll
        for (int e = 0; e < V; ++e) Qs[(kv * V + e) * kQ2 + r] = qreg[i][e];
      }
    }
#pragma unroll
    for (int i = 0; i < NBV; ++i) {
      const int v = tid + i * kT2;
      const int r = v % kB2, kv = v / kB2;
#pragma unroll
      for (int e = 0; e < V; ++e) Bs[(kv * V + e) * kB2 + r] = breg[i][e];
    }
  };
  const int nchunk = (D + kKC - 1) / kKC;
  if (vec) g_load(b_begin, 0);
  for (int64_t b0 = b_begin; b0 < b_end; b0 += kB2) {
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    for (int kc = 0; kc < nchunk; ++kc) {
      const int k0 = kc * kKC;
      // ---- stage the operand tiles k-major (zero padded) ----
      if (vec) {
        s_store();
      } else {
        for (int v = tid; v < kQ2 * kKC; v += kT2) {
          const int r = v % kQ2, k = v / kQ2;
          Qs[k * kQ2 + r] = (q0 + r < n_query && k0 + k < D) ? to_f32(query[(q0 + r) * D + k0 + k]) : 0.f;
        }
        for (int v = tid; v < kB2 * kKC; v += kT2) {
          const int r = v % kB2, k = v / kB2;
          Bs[k * kB2 + r] = (b0 + r < b_end && k0 + k < D) ? to_f32(bank[(b0 + r) * D + k0 + k]) : 0.f;
        }
      }
      __syncthreads();
      if (vec) {
        if (kc + 1 < nchunk) g_load(b0, k0 + kKC);
        else if (b0 + kB2 < b_end) g_load(b0 + kB2, 0);
      }
#pragma unroll 8
      for (int k = 0; k < kKC; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(Qs + k * kQ2 + 4 * ty);
        const float b[4] = {Bs[k * kB2 + tx], Bs[k * kB2 + tx + 32], Bs[k * kB2 + tx + 64], Bs[k * kB2 + tx + 96]};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[0][j] = fmaf(a.x, b[j], acc[0][j]);
          acc[1][j] = fmaf(a.y, b[j], acc[1][j]);
          acc[2][j] = fmaf(a.z, b[j], acc[2][j]);
          acc[3][j] = fmaf(a.w, b[j], acc[3][j]);
        }
      }
      __syncthreads();
    }
    // ---- threshold filter into the candidate buffers ----
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int q = 4 * ty + i;
      if (q0 + q >= n_query) continue;
      const unsigned long long tk = tau[q];
      const int64_t self = exclude_self_offset >= 0 ? exclude_self_offset + q0 + q : -1;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t col = b0 + tx + 32 * j;
        if (col < b_end && col != self) {
          const unsigned long long key = make_key(acc[i][j], (unsigned)col);
          if (key > tk) cand[q * kCap + atomicAdd(&cnt[q], 1)] = key;
        }
      }
    }
    __syncthreads();
    // ---- merge the buffers that could overflow on the next tile (warp w owns queries 4 w .. 4 w + 3) ----
#pragma unroll 1
    for (int i = 0; i < 4; ++i) {
      const int q = 4 * warp + i;
      const int c = cnt[q];
      if (c > kCap - kB2) {
        warp_merge_candidates(run + q * K, cand + q * kCap, c, K, lane);
        if (lane == 0) { cnt[q] = 0; tau[q] = run[q * K + K - 1]; }
      }
    }
    __syncthreads();
  }
#pragma unroll 1
  for (int i = 0; i < 4; ++i) {
    const int q = 4 * warp + i;
    const int c = cnt[q];
    if (c > 0) warp_merge_candidates(run + q * K, cand + q * kCap, c, K, lane);
    __syncwarp();
    if (q0 + q < n_query) {
      unsigned long long* dst = lists + ((int64_t)blockIdx.y * n_query + q0 + q) * K;
      for (int r = lane; r < K; r += 32) dst[r] = run[q * K + r];
    }
  }
}

// per query: merge the per-split lists (each sorted descending), emit values and indices; one warp per query
__global__ void __launch_bounds__(128)
sim_topk_merge_kernel(const unsigned long long* __restrict__ lists, int splits, int64_t n_query, int K, int k,
                      float* __restrict__ vals, int64_t* __restrict__ idx) {
  __shared__ unsigned long long res_all[4 * kMaxK];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t q = (int64_t)blockIdx.x * 4 + warp;
  if (q >= n_query) return;
  unsigned long long* res = res_all + warp * kMaxK;
  for (int i = lane; i < K; i += 32) res[i] = lists[q * K + i];
  for (int s = 1; s < splits; ++s) {
    __syncwarp();
    const unsigned long long* l = lists + ((int64_t)s * n_query + q) * K;
    for (int i = lane; i < K; i += 32) {
      const unsigned long long a = res[i], b = l[K - 1 - i];
      res[i] = a > b ? a : b;
    }
    warp_bitonic_merge_desc(res, K, lane);
  }
  __syncwarp();
  for (int r = lane; r < k; r += 32) {
    const unsigned long long key = res[r];
    const bool ok = key != 0ull;
    vals[q * k + r] = ok ? key_value(key) : -INFINITY;
    idx[q * k + r] = ok ? (int64_t)(~(unsigned)(key & 0xFFFFFFFFull)) : -1;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Third form (sm3_sim_topk_mat): the similarities of a chunk of queries are MATERIALISED (fp32 [chunk, n_bank], at most
// ~256 MB at a time -- what the reference does for the whole matrix) by a register-tiled FP32 contraction that uses every
// SM whatever the query count (64 x 128 tiles, 8 x 4 outputs per thread), then ONE CTA PER QUERY selects the top k with a
// 4-pass radix select on the order-preserving 32-bit keys (warp-aggregated histogram adds), collects the elements above
// the threshold plus the lowest-index ties, and sorts those k.  Work per query is O(n_bank) with no per-tile sorting, and
// the selection parallelises over queries, not over bank splits.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kGQ = 128, kGB = 128, kGK = 16, kGT = 256;

// C[q, b] = sum_k Q[q, k] B[b, k] for a chunk of queries: 128 x 128 tile per CTA, 8 x 8 outputs per thread (rows
// 4 ty .. + 3 and 64 + 4 ty .. + 3, columns 4 tx .. + 3 and 64 + 4 tx .. + 3: every shared-memory read is a conflict-free
// LDS.128 and feeds 16 FMAs), operands staged k-major, the next k-chunk's global loads issued before the FMAs of the
// current one.  fp32 accumulation in ascending k, like the other two forms (identical sums, identical top-k).
template <typename T>
__global__ void __launch_bounds__(kGT)
sim_gemm_kernel(const T* __restrict__ query, const T* __restrict__ bank, int64_t nq, int64_t n_bank, int D, int vec,
                float* __restrict__ out) {
  __shared__ __align__(16) float Qs[kGK * kGQ];
  __shared__ __align__(16) float Bs[kGK * kGB];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t q0 = (int64_t)blockIdx.x * kGQ, b0 = (int64_t)blockIdx.y * kGB;
  constexpr int V = VecIO<T>::N;
  constexpr int NQV = kGQ * kGK / V / kGT, NBV = kGB * kGK / V / kGT;       // 2 / 1 vectors per thread and operand
  static_assert(NQV >= 1 && NBV >= 1, "tile / vector mismatch");
  float qreg[NQV][V], breg[NBV][V];
  auto g_load = [&](int k0) {
#pragma unroll
    for (int i = 0; i < NQV; ++i) {
      const int v = tid + i * kGT;
      const int r = v % kGQ, kv = v / kGQ;
      if (q0 + r < nq && k0 + kv * V < D) VecIO<T>::load(query + (q0 + r) * D + k0 + kv * V, qreg[i]);
      else {
#pragma unroll
        for (int e = 0; e < V; ++e) qreg[i][e] = 0.f;
      }
    }
#pragma unroll
    for (int i = 0; i < NBV; ++i) {
      const int v = tid + i * kGT;
      const int r = v % kGB, kv = v / kGB;
      if (b0 + r < n_bank && k0 + kv * V < D) VecIO<T>::load(bank + (b0 + r) * D + k0 + kv * V, breg[i]);
      else {
#pragma unroll
        for (int e = 0; e < V; ++e) breg[i][e] = 0.f;
      }
    }
  };
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  const int nchunk = (D + kGK - 1) / kGK;
  if (vec) g_load(0);
  for (int kc = 0; kc < nchunk; ++kc) {
    const int k0 = kc * kGK;
    if (vec) {
#pragma unroll
      for (int i = 0; i < NQV; ++i) {
        const int v = tid + i * kGT;
        const int r = v % kGQ, kv = v / kGQ;
#pragma unroll
        for (int e = 0; e < V; ++e) Qs[(kv * V + e) * kGQ + r] = qreg[i][e];
      }
#pragma unroll
      for (int i = 0; i < NBV; ++i) {
        const int v = tid + i * kGT;
        const int r = v % kGB, kv = v / kGB;
#pragma unroll
        for (int e = 0; e < V; ++e) Bs[(kv * V + e) * kGB + r] = breg[i][e];
      }
    } else {
      for (int v = tid; v < kGQ * kGK; v += kGT) {
        const int r = v % kGQ, k = v / kGQ;
        Qs[k * kGQ + r] = (q0 + r < nq && k0 + k < D) ? to_f32(query[(q0 + r) * D + k0 + k]) : 0.f;
      }
      for (int v = tid; v < kGB * kGK; v += kGT) {
        const int r = v % kGB, k = v / kGB;
        Bs[k * kGB + r] = (b0 + r < n_bank && k0 + k < D) ? to_f32(bank[(b0 + r) * D + k0 + k]) : 0.f;
      }
    }
    __syncthreads();
    if (vec && kc + 1 < nchunk) g_load(k0 + kGK);
#pragma unroll
    for (int k = 0; k < kGK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(Qs + k * kGQ + 4 * ty);
      const float4 a1 = *reinterpret_cast<const float4*>(Qs + k * kGQ + 64 + 4 * ty);
      const float4 c0 = *reinterpret_cast<const float4*>(Bs + k * kGB + 4 * tx);
      const float4 c1 = *reinterpret_cast<const float4*>(Bs + k * kGB + 64 + 4 * tx);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  const bool vec_out = (n_bank % 4 == 0);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t q = q0 + (i < 4 ? 4 * ty + i : 64 + 4 * ty + (i - 4));
    if (q >= nq) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t col = b0 + 64 * h + 4 * tx;
      float* dst = out + q * n_bank + col;
      if (vec_out && col + 3 < n_bank) {
        *reinterpret_cast<float4*>(dst) = make_float4(acc[i][4 * h], acc[i][4 * h + 1], acc[i][4 * h + 2], acc[i][4 * h + 3]);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (col + e < n_bank) dst[e] = acc[i][4 * h + e];
      }
    }
  }
}

__device__ __forceinline__ unsigned key32(float v) {
  unsigned u = __float_as_uint(v);
  return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}

// one CTA per query: radix select of the k-th largest key, collection, sort.  sims: this chunk's [nq, n_bank] fp32.
__global__ void __launch_bounds__(kTopkThreads)
topk_radix_kernel(const float* __restrict__ sims, int64_t n_bank, int k, int K, int64_t q_first,
                  int64_t exclude_self_offset, float* __restrict__ vals, int64_t* __restrict__ idx) {
  __shared__ int hist[256];
  __shared__ int wtot[8];
  __shared__ unsigned s_prefix;
  __shared__ int s_need, s_gt, s_eq, s_run, s_found;
  __shared__ unsigned long long sel[kMaxK];
  __shared__ unsigned tie_idx[kMaxK];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t q = blockIdx.x;                                   // query within the chunk
  const float* row = sims + q * n_bank;
  const int64_t self = exclude_self_offset >= 0 ? exclude_self_offset + q_first + q : -1;
  const int64_t n_round = (n_bank + kTopkThreads - 1) / kTopkThreads * kTopkThreads;
  auto key_at = [&](int64_t i) -> unsigned { return (i < n_bank && i != self) ? key32(__ldg(row + i)) : 0u; };
  if (tid == 0) { s_prefix = 0u; s_need = k; s_gt = 0; s_eq = 0; s_run = 0; }
  for (int shift = 24; shift >= 0; shift -= 8) {
    hist[tid] = 0;
    if (tid == 0) s_found = 0;
    __syncthreads();
    const unsigned prefix = s_prefix;
    const unsigned himask = shift == 24 ? 0u : (0xFFFFFFFFu << (shift + 8));
    for (int64_t i = tid; i < n_round; i += kTopkThreads) {
      const unsigned u = key_at(i);
      const bool act = u != 0u && (u & himask) == (prefix & himask);
      const unsigned bin = (u >> shift) & 255u;
      if (shift == 24) {
        // top digit (sign + high exponent bits): a handful of bins take everything -> one add per warp and bin
        const unsigned tag = act ? bin : (256u + (unsigned)lane);   // inactive lanes never group
        const unsigned peers = __match_any_sync(0xffffffffu, tag);
        if (act && lane == __ffs(peers) - 1) atomicAdd(&hist[bin], __popc(peers));
      } else if (act) {
        atomicAdd(&hist[bin], 1);                                   // lower digits are spread: plain shared-memory adds
      }
    }
    __syncthreads();
    // suffix scan from the top bin: thread t looks at bin 255 - t
    const int v = hist[255 - tid];
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) wtot[warp] = incl;
    __syncthreads();
    int off = 0;
    for (int w = 0; w < warp; ++w) off += wtot[w];
    incl += off;
    const int excl = incl - v;
    const int need = s_need;
    __syncthreads();
    if (excl < need && need <= incl) {                              // exactly one bin holds the k-th largest
      s_prefix = prefix | ((unsigned)(255 - tid) << shift);
      s_need = need - excl;
      s_found = 1;
    }
    __syncthreads();
    if (!s_found) {                                                 // fewer than k valid elements: everything is taken
      if (tid == 0) s_prefix = 0u;
      __syncthreads();
      break;
    }
  }
  const unsigned ustar = s_prefix;
  const int need = s_need;                                          // how many elements EQUAL to the threshold are taken
  if (ustar == 0u) {                                                // fewer than k valid elements: take everything valid
    for (int64_t i = tid; i < n_round; i += kTopkThreads) {
      const unsigned u = key_at(i);
      if (u != 0u) { const int s_ = atomicAdd(&s_gt, 1); if (s_ < kMaxK) sel[s_] = ((unsigned long long)u << 32) | (unsigned long long)(~(unsigned)i); }
    }
    __syncthreads();
  } else {
    for (int64_t i = tid; i < n_round; i += kTopkThreads) {
      const unsigned u = key_at(i);
      if (u > ustar) sel[atomicAdd(&s_gt, 1)] = ((unsigned long long)u << 32) | (unsigned long long)(~(unsigned)i);
      else if (u == ustar) { const int e = atomicAdd(&s_eq, 1); if (e < kMaxK) tie_idx[e] = (unsigned)i; }
    }
    __syncthreads();
    const int gt = s_gt, eq = s_eq;                                 // gt == k - need by construction
    if (eq == need) {                                               // every tie is taken: no ordering needed
      if (tid < need) sel[gt + tid] = ((unsigned long long)ustar << 32) | (unsigned long long)(~tie_idx[tid]);
    } else {
      // more ties than needed: the lowest bank indices win -> ordered scan over the row
      for (int64_t base = 0; base < n_round && s_run < need; base += kTopkThreads) {
        const bool f = key_at(base + tid) == ustar;
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        if (lane == 0) wtot[warp] = __popc(bal);
        __syncthreads();
        int before = s_run;
        for (int w = 0; w < warp; ++w) before += wtot[w];
        const int pos = before + __popc(bal & ((1u << lane) - 1u));
        if (f && pos < need) sel[gt + pos] = ((unsigned long long)ustar << 32) | (unsigned long long)(~(unsigned)(base + tid));
        __syncthreads();
        if (tid == 0) { int t = 0; for (int w = 0; w < 8; ++w) t += wtot[w]; s_run += t; }
        __syncthreads();
      }
    }
    __syncthreads();
  }
  const int have = min(k, ustar == 0u ? min(s_gt, kMaxK) : k);
  for (int i = tid; i < K; i += kTopkThreads)
    if (i >= have) sel[i] = 0ull;
  bitonic_sort_desc(sel, K);
  const int64_t qg = q_first + q;
  for (int r = tid; r < k; r += kTopkThreads) {
    const unsigned long long key = sel[r];
    const bool ok = key != 0ull;
    vals[qg * k + r] = ok ? key_value(key) : -INFINITY;
    idx[qg * k + r] = ok ? (int64_t)(~(unsigned)(key & 0xFFFFFFFFull)) : -1;
  }
}

struct Topk3Plan { int K; int64_t chunk; size_t ws; };
Topk3Plan topk3_plan(int64_t n_query, int64_t n_bank, int k) {
  Topk3Plan p{};
  p.K = 8;
  while (p.K < k) p.K <<= 1;
  int64_t chunk = ((int64_t)256 << 20) / (n_bank * 4);
  chunk = chunk / kGQ * kGQ;
  if (chunk < kGQ) chunk = kGQ;
  if (chunk > n_query) chunk = n_query;
  p.chunk = chunk;
  p.ws = (size_t)chunk * n_bank * sizeof(float);
  return p;
}

struct Topk2Plan { int K, splits; int64_t rows_per_split; size_t smem, ws; };
Topk2Plan topk2_plan(int64_t n_query, int64_t n_bank, int k) {
  Topk2Plan p{};
  p.K = 8;
  while (p.K < k) p.K <<= 1;
  const int64_t qt = (n_query + kQ2 - 1) / kQ2;
  int64_t want = num_sms() / qt;
  if (want < 1) want = 1;
  if (want > 64) want = 64;
  int64_t rps = (n_bank + want - 1) / want;
  rps = (rps + kB2 - 1) / kB2 * kB2;
  p.rows_per_split = rps;
  p.splits = (int)((n_bank + rps - 1) / rps);
  p.smem = (size_t)(kKC * kQ2 + kKC * kB2) * 4 + (size_t)kQ2 * p.K * 8 + (size_t)kQ2 * kCap * 8 + kQ2 * 8 + kQ2 * 4;
  p.ws = (size_t)p.splits * n_query * p.K * 8;
  return p;
}

}  // namespace
}  // namespace sm3

using namespace sm3;

extern "C" int sm3_sim_topk(const void* query, const void* bank, int64_t n_query, int64_t n_bank, int D, int dtype,
                            int k, int64_t exclude_self_offset, float* vals, int64_t* idx, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SM3_REQUIRE(query && bank && vals && idx, SM3_ERR_SHAPE, "sim_topk: null pointer");
  SM3_REQUIRE(dtype_ok(dtype), SM3_ERR_DTYPE, "sim_topk: bad dtype %d", dtype);
  SM3_REQUIRE(n_query >= 1 && n_bank >= 1 && D >= 1 && D <= 4096, SM3_ERR_SHAPE, "sim_topk: bad shape");
  SM3_REQUIRE(n_bank < ((int64_t)1 << 32) - 1, SM3_ERR_SHAPE, "sim_topk: bank too large");
  SM3_REQUIRE(k >= 1 && k <= kMaxK && k <= n_bank, SM3_ERR_SHAPE, "sim_topk: need 1 <= k <= min(%d, n_bank), got %d",
              kMaxK, k);
  int K = 1;
  while (K < k) K <<= 1;
  const size_t smem = (size_t)kQT * K * 8 + (size_t)kCH * 8 + (size_t)kQT * kCH * 4 + (size_t)kQT * D * 4;
  const unsigned grid = (unsigned)((n_query + kQT - 1) / kQT);
  SM3_DISPATCH_DTYPE(dtype, T, {
    SM3_CHECK_CUDA(cudaFuncSetAttribute(sim_topk_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sim_topk_kernel<T><<<grid, kTopkThreads, smem, st>>>((const T*)query, (const T*)bank, n_query, n_bank, D, k, K,
                                                         exclude_self_offset, vals, idx);
  });
  SM3_CHECK_CUDA(cudaGetLastError());
  return SM3_OK;
}

extern "C" size_t sm3_sim_topk_workspace_bytes(int64_t n_query, int64_t n_bank, int k) {
  if (n_query < 1 || n_bank < 1 || k < 1 || k > kMaxK) return 0;
  return topk2_plan(n_query, n_bank, k).ws;
}

extern "C" int sm3_sim_topk_ws(const void* query, const void* bank, int64_t n_query, int64_t n_bank, int D, int dtype,
                               int k, int64_t exclude_self_offset, float* vals, int64_t* idx, void* workspace,
                               size_t workspace_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SM3_REQUIRE(query && bank && vals && idx && workspace, SM3_ERR_SHAPE, "sim_topk_ws: null pointer");
  SM3_REQUIRE(dtype_ok(dtype), SM3_ERR_DTYPE, "sim_topk_ws: bad dtype %d", dtype);
  SM3_REQUIRE(n_query >= 1 && n_bank >= 1 && D >= 1, SM3_ERR_SHAPE, "sim_topk_ws: bad shape");
  SM3_REQUIRE(n_bank < ((int64_t)1 << 32) - 1, SM3_ERR_SHAPE, "sim_topk_ws: bank too large");
  SM3_REQUIRE(k >= 1 && k <= kMaxK && k <= n_bank, SM3_ERR_SHAPE, "sim_topk_ws: need 1 <= k <= min(%d, n_bank), got %d",
              kMaxK, k);
  const Topk2Plan pl = topk2_plan(n_query, n_bank, k);
  SM3_REQUIRE(workspace_bytes >= pl.ws, SM3_ERR_WORKSPACE, "sim_topk_ws: workspace %zu < %zu", workspace_bytes, pl.ws);
  SM3_REQUIRE((((uintptr_t)workspace) & 7u) == 0, SM3_ERR_SHAPE, "sim_topk_ws: workspace must be 8-byte aligned");
  unsigned long long* lists = (unsigned long long*)workspace;
  const dim3 grid((unsigned)((n_query + kQ2 - 1) / kQ2), (unsigned)pl.splits);
  SM3_DISPATCH_DTYPE(dtype, T, {
    const int vec = (D % VecIO<T>::N == 0) && ((((uintptr_t)bank) & 15u) == 0) && ((((uintptr_t)query) & 15u) == 0);
    SM3_CHECK_CUDA(cudaFuncSetAttribute(sim_topk_tile_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    sim_topk_tile_kernel<T><<<grid, kT2, pl.smem, st>>>((const T*)query, (const T*)bank, n_query, n_bank, D, pl.K,
                                                        exclude_self_offset, pl.rows_per_split, vec, lists);
  });
  SM3_CHECK_CUDA(cudaGetLastError());
  sim_topk_merge_kernel<<<(unsigned)((n_query + 3) / 4), 128, 0, st>>>(lists, pl.splits, n_query, pl.K, k, vals, idx);
  SM3_CHECK_CUDA(cudaGetLastError());
  return SM3_OK;
}

extern "C" size_t sm3_sim_topk_mat_workspace_bytes(int64_t n_query, int64_t n_bank, int k) {
  if (n_query < 1 || n_bank < 1 || k < 1 || k > kMaxK) return 0;
  return topk3_plan(n_query, n_bank, k).ws;
}

extern "C" int sm3_sim_topk_mat(const void* query, const void* bank, int64_t n_query, int64_t n_bank, int D, int dtype,
                                int k, int64_t exclude_self_offset, float* vals, int64_t* idx, void* workspace,
                                size_t workspace_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SM3_REQUIRE(query && bank && vals && idx && workspace, SM3_ERR_SHAPE, "sim_topk_mat: null pointer");
  SM3_REQUIRE(dtype_ok(dtype), SM3_ERR_DTYPE, "sim_topk_mat: bad dtype %d", dtype);
  SM3_REQUIRE(n_query >= 1 && n_bank >= 1 && D >= 1, SM3_ERR_SHAPE, "sim_topk_mat: bad shape");
  SM3_REQUIRE(n_bank < ((int64_t)1 << 31), SM3_ERR_SHAPE, "sim_topk_mat: bank too large");
  SM3_REQUIRE(k >= 1 && k <= kMaxK && k <= n_bank, SM3_ERR_SHAPE, "sim_topk_mat: need 1 <= k <= min(%d, n_bank), got %d",
              kMaxK, k);
  const Topk3Plan pl = topk3_plan(n_query, n_bank, k);
  SM3_REQUIRE(workspace_bytes >= pl.ws, SM3_ERR_WORKSPACE, "sim_topk_mat: workspace %zu < %zu", workspace_bytes, pl.ws);
  SM3_REQUIRE((((uintptr_t)workspace) & 15u) == 0, SM3_ERR_SHAPE, "sim_topk_mat: workspace must be 16-byte aligned");
  float* sims = (float*)workspace;
  const size_t esz = (size_t)dtype_size(dtype);
  for (int64_t q0 = 0; q0 < n_query; q0 += pl.chunk) {
    const int64_t nq = n_query - q0 < pl.chunk ? n_query - q0 : pl.chunk;
    const dim3 grid((unsigned)((nq + kGQ - 1) / kGQ), (unsigned)((n_bank + kGB - 1) / kGB));
    const char* qptr = (const char*)query + (size_t)q0 * D * esz;
    SM3_DISPATCH_DTYPE(dtype, T, {
      const int vec = (D % VecIO<T>::N == 0) && ((((uintptr_t)bank) & 15u) == 0) && ((((uintptr_t)qptr) & 15u) == 0);
      sim_gemm_kernel<T><<<grid, kGT, 0, st>>>((const T*)qptr, (const T*)bank, nq, n_bank, D, vec, sims);
    });
    SM3_CHECK_CUDA(cudaGetLastError());
    topk_radix_kernel<<<(unsigned)nq, kTopkThreads, 0, st>>>(sims, n_bank, k, pl.K, q0, exclude_self_offset, vals, idx);
    SM3_CHECK_CUDA(cudaGetLastError());
  }
  return SM3_OK;
}
