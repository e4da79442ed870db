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
