// N4: DeepCluster spherical k-means of the memory bank -- the rank-0 loop of `cluster_memory`
// (reference tools/mlc_train.py:144-176) as two kernels per iteration instead of a cuBLAS GEMM, a `.cpu().numpy()` round
// trip, scipy.sparse and a Python loop over the clusters (:153-172):
//
//   kmeans_assign_kernel   E step + the M step's sums in ONE pass over the bank (HBM-bound: n * D * 4 bytes read once):
//                          one warp per sample keeps the row in registers, takes its dot product with the K centroids
//                          (shared memory), writes the argmax (ties -> lower centroid, as `dot_products.max(dim=1)`), and
//                          adds the row to its own per-cluster register accumulators; warps are folded in a fixed order
//                          -> per-CTA partial sums / counts -> one more fixed-order fold (last CTA): deterministic,
//                          no float atomics.
//   kmeans_update_kernel   centroid_k = sum_k / count_k for non-empty clusters (empty ones keep the old centroid, :173),
//                          then L2 normalisation of ALL centroids (:175).
// Between the two a multi-rank caller all-reduces (sums, counts): every rank then holds identical centroids while the
// bank stays sharded (the reference gathers the whole bank to rank 0 and broadcasts the result, :137-143, :185-186).
//
// Supported shape for this fused form: fp32 bank, D % 128 == 0, D <= 512, K <= 8 (the reference's heads have
// K in {2, 3, 5} and D = mlc_proj_dim in {256, 512}); other shapes use sm3_sim_topk + a one-hot GEMM from Python.
#include "common.cuh"

namespace sm3 {
namespace {

constexpr int kKmThreads = 256;
constexpr int kKmWarps = kKmThreads / 32;
constexpr int kKmMaxK = 8;

// smem: centroids [K][D] | fold buffer [K][D] | warp counts [warps][K]
template <int CH>   // CH = D / 128: 16-byte chunks per lane
__global__ void __launch_bounds__(kKmThreads)
kmeans_assign_kernel(const float* __restrict__ emb, int64_t n, int K, const float* __restrict__ cent,
                     int64_t* __restrict__ assign, float* __restrict__ part_sums, float* __restrict__ part_counts,
                     float* __restrict__ sums, float* __restrict__ counts, unsigned* __restrict__ ticket,
                     int64_t rows_per_cta) {
  constexpr int D = 128 * CH;
  extern __shared__ float km_smem[];
  float* sc = km_smem;                       // [K][D]
  float* fold = sc + kKmMaxK * D;            // [K][D]
  float* wcnt = fold + kKmMaxK * D;          // [warps][kKmMaxK]
  __shared__ bool is_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < K * D; i += kKmThreads) sc[i] = cent[i];
  for (int i = tid; i < kKmMaxK * D; i += kKmThreads) fold[i] = 0.f;
  __syncthreads();

  float acc[kKmMaxK][CH][4];
  float cnt[kKmMaxK];
#pragma unroll
  for (int k = 0; k < kKmMaxK; ++k) {
    cnt[k] = 0.f;
#pragma unroll
    for (int c = 0; c < CH; ++c)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[k][c][e] = 0.f;
  }
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_cta;
  const int64_t r_end = min(n, r_begin + rows_per_cta);
  for (int64_t r = r_begin + warp; r < r_end; r += kKmWarps) {
    float x[CH][4];
#pragma unroll
    for (int c = 0; c < CH; ++c) VecIO<float>::load(emb + r * D + (lane + 32 * c) * 4, x[c]);
    int best = 0;
    float best_v = -INFINITY;
#pragma unroll
    for (int k = 0; k < kKmMaxK; ++k) {
      if (k < K) {
        float a = 0.f;
#pragma unroll
        for (int c = 0; c < CH; ++c) {
          const float4 cv = *reinterpret_cast<const float4*>(sc + k * D + (lane + 32 * c) * 4);
          a = fmaf(cv.x, x[c][0], a); a = fmaf(cv.y, x[c][1], a); a = fmaf(cv.z, x[c][2], a); a = fmaf(cv.w, x[c][3], a);
        }
        a = warp_sum(a);
        if (a > best_v) { best_v = a; best = k; }          // strict: ties keep the lower centroid index
      }
    }
    if (lane == 0) assign[r] = best;
#pragma unroll
    for (int k = 0; k < kKmMaxK; ++k) {
      const bool mine = (k == best);
      cnt[k] += mine ? 1.f : 0.f;
#pragma unroll
      for (int c = 0; c < CH; ++c)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[k][c][e] += mine ? x[c][e] : 0.f;
    }
  }
  if (part_sums == nullptr) return;           // E step only (the final assignment pass)
  // ---- fold the warps in a fixed order ----
  for (int w = 0; w < kKmWarps; ++w) {
    if (warp == w) {
#pragma unroll
      for (int k = 0; k < kKmMaxK; ++k) {
        if (k < K) {
#pragma unroll
          for (int c = 0; c < CH; ++c) {
            float4* dst = reinterpret_cast<float4*>(fold + k * D + (lane + 32 * c) * 4);
            float4 v = *dst;
            v.x += acc[k][c][0]; v.y += acc[k][c][1]; v.z += acc[k][c][2]; v.w += acc[k][c][3];
            *dst = v;
          }
        }
      }
      if (lane == 0)
        for (int k = 0; k < kKmMaxK; ++k) wcnt[w * kKmMaxK + k] = cnt[k];
    }
    __syncthreads();
  }
  float* ps = part_sums + (size_t)blockIdx.x * K * D;
  for (int i = tid; i < K * D; i += kKmThreads) ps[i] = fold[i];
  if (tid < K) {
    float c = 0.f;
    for (int w = 0; w < kKmWarps; ++w) c += wcnt[w * kKmMaxK + tid];
    part_counts[(size_t)blockIdx.x * K + tid] = c;
  }
  // ---- the CTA that finishes last folds the per-CTA partials, again in a fixed order ----
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const unsigned t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int i = tid; i < K * D; i += kKmThreads) {
    float s = 0.f;
    for (unsigned b = 0; b < gridDim.x; ++b) s += __ldcg(part_sums + (size_t)b * K * D + i);
    sums[i] = s;
  }
  if (tid < K) {
    float c = 0.f;
    for (unsigned b = 0; b < gridDim.x; ++b) c += __ldcg(part_counts + (size_t)b * K + tid);
    counts[tid] = c;
  }
  if (tid == 0) *ticket = 0u;
}

// one CTA per centroid
__global__ void __launch_bounds__(128)
kmeans_update_kernel(const float* __restrict__ sums, const float* __restrict__ counts, const float* __restrict__ cent_old,
                     float* __restrict__ cent_new, int D, float eps) {
  __shared__ float red[32];
  const int k = blockIdx.x;
  const float c = counts[k];
  float ss = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float v = c > 0.f ? sums[(size_t)k * D + d] / c : cent_old[(size_t)k * D + d];
    ss = fmaf(v, v, ss);
  }
  ss = block_sum(ss, red);
  const float nrm = fmaxf(sqrtf(ss), eps);
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float v = c > 0.f ? sums[(size_t)k * D + d] / c : cent_old[(size_t)k * D + d];
    cent_new[(size_t)k * D + d] = v / nrm;
  }
}

int64_t km_rows_per_cta(int64_t n, int* grid) {
  const int64_t cap = (int64_t)num_sms() * 4;                     // <= 4 resident CTAs per SM is plenty for an HBM-bound pass
  int64_t g = (n + kKmWarps * 4 - 1) / (kKmWarps * 4);             // at least 4 rows per warp
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  const int64_t rpc = (n + g - 1) / g;
  *grid = (int)((n + rpc - 1) / rpc);
  return rpc;
}

}  // namespace
}  // namespace sm3

using namespace sm3;

extern "C" int sm3_kmeans_supported(int D, int K, int dtype) {
  return (dtype == SM3_F32 && D % 128 == 0 && D >= 128 && D <= 512 && K >= 1 && K <= kKmMaxK) ? 1 : 0;
}

extern "C" size_t sm3_kmeans_workspace_bytes(int64_t n, int D, int K) {
  int grid = 1;
  km_rows_per_cta(n > 0 ? n : 1, &grid);
  return ((size_t)grid * K * D + (size_t)grid * K) * sizeof(float) + 256;
}

extern "C" int sm3_kmeans_assign(const float* emb, int64_t n, int D, const float* centroids, int K, int64_t* assign,
                                 float* sums, float* counts, void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SM3_REQUIRE(emb && centroids && assign, SM3_ERR_SHAPE, "kmeans_assign: null pointer");
  SM3_REQUIRE(n >= 1, SM3_ERR_SHAPE, "kmeans_assign: empty bank");
  SM3_REQUIRE(sm3_kmeans_supported(D, K, SM3_F32), SM3_ERR_DTYPE,
              "kmeans_assign: needs D %% 128 == 0, D <= 512, K <= %d (got D=%d K=%d)", kKmMaxK, D, K);
  SM3_REQUIRE(aligned16(emb) && aligned16(centroids), SM3_ERR_SHAPE, "kmeans_assign: rows must be 16-byte aligned");
  SM3_REQUIRE((sums == nullptr) == (counts == nullptr), SM3_ERR_SHAPE, "kmeans_assign: sums/counts must both be given or both NULL");
  int grid = 1;
  const int64_t rpc = km_rows_per_cta(n, &grid);
  float *ps = nullptr, *pc = nullptr;
  unsigned* ticket = nullptr;
  if (sums != nullptr) {
    SM3_REQUIRE(workspace && workspace_bytes >= sm3_kmeans_workspace_bytes(n, D, K), SM3_ERR_WORKSPACE,
                "kmeans_assign: workspace too small");
    ps = (float*)workspace;
    pc = ps + (size_t)grid * K * D;
    ticket = (unsigned*)(pc + (size_t)grid * K);
    SM3_CHECK_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned), st));
  }
  const size_t smem = ((size_t)2 * kKmMaxK * D + kKmWarps * kKmMaxK) * sizeof(float);
#define SM3_KM_LAUNCH(CH)                                                                                              \
  do {                                                                                                                 \
    SM3_CHECK_CUDA(cudaFuncSetAttribute(kmeans_assign_kernel<CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    kmeans_assign_kernel<CH><<<grid, kKmThreads, smem, st>>>(emb, n, K, centroids, assign, ps, pc, sums, counts, ticket, rpc); \
  } while (0)
  switch (D / 128) {
    case 1: SM3_KM_LAUNCH(1); break;
    case 2: SM3_KM_LAUNCH(2); break;
    case 3: SM3_KM_LAUNCH(3); break;
    default: SM3_KM_LAUNCH(4); break;
  }
#undef SM3_KM_LAUNCH
  SM3_CHECK_CUDA(cudaGetLastError());
  return SM3_OK;
}

extern "C" int sm3_kmeans_update(const float* sums, const float* counts, const float* centroids_old, float* centroids_new,
                                 int D, int K, float eps, void* stream) {
  SM3_REQUIRE(sums && counts && centroids_old && centroids_new, SM3_ERR_SHAPE, "kmeans_update: null pointer");
  SM3_REQUIRE(D >= 1 && K >= 1 && eps > 0.f, SM3_ERR_SHAPE, "kmeans_update: bad shape");
  kmeans_update_kernel<<<K, 128, 0, (cudaStream_t)stream>>>(sums, counts, centroids_old, centroids_new, D, eps);
  SM3_CHECK_CUDA(cudaGetLastError());
  return SM3_OK;
}
