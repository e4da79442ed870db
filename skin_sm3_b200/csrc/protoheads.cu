// N3 (tail of the multi-label head block): optional L2 normalisation of the self-attention features + the eight
// bias-free prototype Linears of `Model.forward` (reference tools/mlc_train.py:81-87)
//
//     for i in range(len(sa_feats)): sa_feats[i] = F.normalize(sa_feats[i], dim=-1)        # if l2_norm
//     preds = [prototypes[i](sa_feats[i % len(sa_feats)]) for i in range(8)]
//
// as ONE launch forward and ONE backward (plus a cuBLAS GEMM for dW), instead of 8 normalise chains + 8 small GEMMs and
// their ~40 backward launches: one warp per feature row keeps the row in registers, normalises it, and takes its dot
// products with the prototype rows of the heads that read it (all C = 24 prototype rows live in shared memory); the
// [B, 24] logits come out in the layout the fused 8-head CE (K4, heads.cu) consumes.  HBM-bound: the [Hf, B, D] features
// are read once (and written once when normalised).  Backward: d feats = (sum_c dlogit_c W_c) through the normalisation.
#include "common.cuh"

namespace sm3 {
namespace {

constexpr int kPhThreads = 256;
constexpr int kPhMaxC = 64;
struct ProtoMap {
  int C;
  unsigned char slot[kPhMaxC];   // feature slot (head % Hf) every class reads
};

template <typename T, int NV>   // NV = D / (32 * VecIO<T>::N) vectors per lane
__global__ void __launch_bounds__(kPhThreads)
proto_heads_fwd_kernel(const T* __restrict__ feats, int Hf, int64_t B, const float* __restrict__ W, ProtoMap map,
                       int l2_norm, float eps, T* __restrict__ z_out, float* __restrict__ inv_norm,
                       float* __restrict__ logits) {
  constexpr int V = VecIO<T>::N;
  constexpr int D = NV * 32 * V;
  extern __shared__ float ph_smem[];   // [C][D]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < map.C * D; i += kPhThreads) ph_smem[i] = W[i];
  __syncthreads();
  const int64_t rows = (int64_t)Hf * B;
  for (int64_t r = (int64_t)blockIdx.x * (kPhThreads / 32) + warp; r < rows; r += (int64_t)gridDim.x * (kPhThreads / 32)) {
    const int s = (int)(r / B);
    const int64_t b = r - (int64_t)s * B;
    float x[NV][V];
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      VecIO<T>::load(feats + r * D + (c * 32 + lane) * V, x[c]);
#pragma unroll
      for (int e = 0; e < V; ++e) ss = fmaf(x[c][e], x[c][e], ss);
    }
    if (l2_norm) {
      ss = warp_sum(ss);
      const float inv = 1.0f / fmaxf(sqrtf(ss), eps);
#pragma unroll
      for (int c = 0; c < NV; ++c) {
#pragma unroll
        for (int e = 0; e < V; ++e) x[c][e] *= inv;
        VecIO<T>::store(z_out + r * D + (c * 32 + lane) * V, x[c]);
        if (sizeof(T) < 4) {       // the logits see the values the caller gets back (rounded to T)
          float y[V];
          VecIO<T>::load(z_out + r * D + (c * 32 + lane) * V, y);
#pragma unroll
          for (int e = 0; e < V; ++e) x[c][e] = y[e];
        }
      }
      if (lane == 0) inv_norm[r] = inv;
    }
    for (int cls = 0; cls < map.C; ++cls) {
      if (map.slot[cls] != s) continue;
      const float* w = ph_smem + cls * D;
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < NV; ++c)
#pragma unroll
        for (int e = 0; e < V; ++e) acc = fmaf(x[c][e], w[(c * 32 + lane) * V + e], acc);
      acc = warp_sum(acc);
      if (lane == 0) logits[b * map.C + cls] = acc;
    }
  }
}

template <typename T, int NV>
__global__ void __launch_bounds__(kPhThreads)
proto_heads_bwd_kernel(const T* __restrict__ z, int Hf, int64_t B, const float* __restrict__ W, ProtoMap map, int l2_norm,
                       const float* __restrict__ inv_norm, const float* __restrict__ dlogits,
                       const T* __restrict__ d_extra, T* __restrict__ d_feats) {
  constexpr int V = VecIO<T>::N;
  constexpr int D = NV * 32 * V;
  extern __shared__ float ph_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < map.C * D; i += kPhThreads) ph_smem[i] = W[i];
  __syncthreads();
  const int64_t rows = (int64_t)Hf * B;
  for (int64_t r = (int64_t)blockIdx.x * (kPhThreads / 32) + warp; r < rows; r += (int64_t)gridDim.x * (kPhThreads / 32)) {
    const int s = (int)(r / B);
    const int64_t b = r - (int64_t)s * B;
    float dz[NV][V];
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      if (d_extra != nullptr) VecIO<T>::load(d_extra + r * D + (c * 32 + lane) * V, dz[c]);
      else {
#pragma unroll
        for (int e = 0; e < V; ++e) dz[c][e] = 0.f;
      }
    }
    for (int cls = 0; cls < map.C; ++cls) {
      if (map.slot[cls] != s) continue;
      const float g = dlogits[b * map.C + cls];
      const float* w = ph_smem + cls * D;
#pragma unroll
      for (int c = 0; c < NV; ++c)
#pragma unroll
        for (int e = 0; e < V; ++e) dz[c][e] = fmaf(g, w[(c * 32 + lane) * V + e], dz[c][e]);
    }
    if (l2_norm) {
      float zz[NV][V];
      float dot = 0.f;
#pragma unroll
      for (int c = 0; c < NV; ++c) {
        VecIO<T>::load(z + r * D + (c * 32 + lane) * V, zz[c]);
#pragma unroll
        for (int e = 0; e < V; ++e) dot = fmaf(zz[c][e], dz[c][e], dot);
      }
      dot = warp_sum(dot);
      const float inv = inv_norm[r];
#pragma unroll
      for (int c = 0; c < NV; ++c)
#pragma unroll
        for (int e = 0; e < V; ++e) dz[c][e] = (dz[c][e] - zz[c][e] * dot) * inv;
    }
#pragma unroll
    for (int c = 0; c < NV; ++c) VecIO<T>::store(d_feats + r * D + (c * 32 + lane) * V, dz[c]);
  }
}

int fill_map(ProtoMap& m, int C, const int* class_slot_host, int Hf) {
  SM3_REQUIRE(C >= 1 && C <= kPhMaxC && class_slot_host != nullptr, SM3_ERR_SHAPE, "proto_heads: need 1 <= C <= %d", kPhMaxC);
  m.C = C;
  for (int i = 0; i < C; ++i) {
    SM3_REQUIRE(class_slot_host[i] >= 0 && class_slot_host[i] < Hf, SM3_ERR_SHAPE, "proto_heads: class %d reads slot %d of %d",
                i, class_slot_host[i], Hf);
    m.slot[i] = (unsigned char)class_slot_host[i];
  }
  return SM3_OK;
}
bool ph_shape_ok(int D, int C, int dtype) {
  const int v = dtype == SM3_F32 ? 4 : 8;
  return dtype_ok(dtype) && D >= 32 * v && D % (32 * v) == 0 && D / (32 * v) <= 4 && C >= 1 && C <= kPhMaxC &&
         (size_t)C * D * 4 <= 160 * 1024;
}

}  // namespace
}  // namespace sm3

using namespace sm3;

extern "C" int sm3_proto_heads_supported(int D, int C, int dtype) { return ph_shape_ok(D, C, dtype) ? 1 : 0; }

#define SM3_PH_DISPATCH(KERNEL, ...)                                                                         \
  do {                                                                                                       \
    SM3_DISPATCH_DTYPE(dtype, T, {                                                                           \
      const int nv = D / (32 * VecIO<T>::N);                                                                 \
      auto go = [&](auto tag) {                                                                              \
        constexpr int NV = decltype(tag)::value;                                                             \
        SM3_CHECK_CUDA(cudaFuncSetAttribute(KERNEL<T, NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        KERNEL<T, NV><<<grid, kPhThreads, smem, st>>>(__VA_ARGS__);                                          \
        return SM3_OK;                                                                                       \
      };                                                                                                     \
      int rc_ = nv == 1 ? go(std::integral_constant<int, 1>{}) : nv == 2 ? go(std::integral_constant<int, 2>{})  \
              : nv == 3 ? go(std::integral_constant<int, 3>{}) : go(std::integral_constant<int, 4>{});       \
      if (rc_) return rc_;                                                                                   \
    });                                                                                                      \
  } while (0)

extern "C" int sm3_proto_heads_fwd(const void* feats, int dtype, int Hf, int64_t B, int D, const float* W_cat, int C,
                                   const int* class_slot_host, int l2_norm, float eps, void* z_out, float* inv_norm,
                                   float* logits, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SM3_REQUIRE(feats && W_cat && logits, SM3_ERR_SHAPE, "proto_heads_fwd: null pointer");
  SM3_REQUIRE(!l2_norm || (z_out && inv_norm), SM3_ERR_SHAPE, "proto_heads_fwd: l2_norm needs z_out and inv_norm");
  SM3_REQUIRE(Hf >= 1 && Hf <= 255 && B >= 1, SM3_ERR_SHAPE, "proto_heads_fwd: bad shape");
  SM3_REQUIRE(ph_shape_ok(D, C, dtype), SM3_ERR_SHAPE, "proto_heads_fwd: unsupported D=%d C=%d dtype=%d", D, C, dtype);
  ProtoMap map;
  int rc = fill_map(map, C, class_slot_host, Hf);
  if (rc) return rc;
  const size_t smem = (size_t)C * D * sizeof(float);
  const int64_t rows = (int64_t)Hf * B;
  int64_t want = (rows + kPhThreads / 32 - 1) / (kPhThreads / 32);
  const unsigned grid = (unsigned)(want < 2 * num_sms() ? (want < 1 ? 1 : want) : 2 * num_sms());
  SM3_PH_DISPATCH(proto_heads_fwd_kernel, (const T*)feats, Hf, B, W_cat, map, l2_norm, eps, (T*)z_out, inv_norm, logits);
  SM3_CHECK_CUDA(cudaGetLastError());
  return SM3_OK;
}

extern "C" int sm3_proto_heads_bwd(const void* z_or_feats, int dtype, int Hf, int64_t B, int D, const float* W_cat, int C,
                                   const int* class_slot_host, int l2_norm, const float* inv_norm, const float* dlogits,
                                   const void* d_extra, void* d_feats, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SM3_REQUIRE(W_cat && dlogits && d_feats, SM3_ERR_SHAPE, "proto_heads_bwd: null pointer");
  SM3_REQUIRE(!l2_norm || (z_or_feats && inv_norm), SM3_ERR_SHAPE, "proto_heads_bwd: l2_norm needs the normalised rows and inv_norm");
  SM3_REQUIRE(Hf >= 1 && Hf <= 255 && B >= 1, SM3_ERR_SHAPE, "proto_heads_bwd: bad shape");
  SM3_REQUIRE(ph_shape_ok(D, C, dtype), SM3_ERR_SHAPE, "proto_heads_bwd: unsupported D=%d C=%d dtype=%d", D, C, dtype);
  ProtoMap map;
  int rc = fill_map(map, C, class_slot_host, Hf);
  if (rc) return rc;
  const size_t smem = (size_t)C * D * sizeof(float);
  const int64_t rows = (int64_t)Hf * B;
  int64_t want = (rows + kPhThreads / 32 - 1) / (kPhThreads / 32);
  const unsigned grid = (unsigned)(want < 2 * num_sms() ? (want < 1 ? 1 : want) : 2 * num_sms());
  SM3_PH_DISPATCH(proto_heads_bwd_kernel, (const T*)z_or_feats, Hf, B, W_cat, map, l2_norm, inv_norm, dlogits,
                  (const T*)d_extra, (T*)d_feats);
  SM3_CHECK_CUDA(cudaGetLastError());
  return SM3_OK;
}
