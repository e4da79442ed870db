// N2: fused projector tail -- the last two layers of `make_projector` (reference src/models/simclr.py:25-26:
// nn.Linear(in_dim, proj_dim, bias=False) -> nn.BatchNorm1d(proj_dim, affine=False)) fused with the F.normalize that
// follows them (:62, :138, :294), so that the kernels of K2/K3 are fed bf16 unit rows directly:
//
//   forward   Y = H W^T                                  proj_gemm_kernel   tcgen05 GEMM [R, K] x [K, D], fp32 in TMEM;
//             (sum_r Y, sum_r Y^2) per column            epilogue           warp transpose-reduce, fixed-order folds
//             [ all-reduce of 2 D + 1 floats under SyncBatchNorm, tools/backbone_train.py:510 ]
//             y^ = (Y - mean) rstd ; z = y^ / max(|y^|, eps) proj_bn_l2_kernel   -> bf16 z, fp32 inv_norm; running stats
//   backward  dy^ = inv (dz - z <z, dz>) ; column sums of dy^ and dy^ * y^      proj_bwd1_kernel
//             [ all-reduce of 2 D floats under SyncBatchNorm ]
//             dY = rstd (dy^ - mean_r dy^ - y^ mean_r (dy^ y^))                 proj_bwd2_kernel -> dY in H's dtype
//             dW = dY^T H, dH = dY W are plain library GEMMs (cuBLAS through torch.matmul) in the Python front end.
//
// Two [R, D] round trips of the stock sequence (BatchNorm's statistics pass and its apply pass) plus K1 disappear; the
// bytes that remain are H (R K b) read once -- the GEMM is HBM-bound at these shapes (K = 2048, D <= 256).
#include "common.cuh"
#include "ptx_sm100.cuh"

namespace sm3 {
namespace {

using namespace ptx;

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_tmap_16bit(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, bool half) {
  static EncodeTiledFn2 fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn2)p;
  }
  SM3_REQUIRE(fn != nullptr, SM3_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2,
                  const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SM3_REQUIRE(r == CUDA_SUCCESS, SM3_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu box_rows=%u", (int)r,
              (unsigned long long)rows, (unsigned long long)cols, box_rows);
  return SM3_OK;
}

// instruction descriptor for kind::f16 with fp16 (0) or bf16 (1) operands, fp32 accumulate, both operands K-major
__host__ __device__ constexpr uint32_t idesc_16bit(int M, int N, int fmt) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// lane L of the warp ends with the sum over all 32 lanes of v[L] (31 shuffles instead of 160)
__device__ __forceinline__ float warp_transpose_reduce(float (&v)[32]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? v[i] : v[i + off];
      const float keep = up ? v[i + off] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return v[0];
}

struct GemmParams {
  int R, K, D;
  float* y;              // [R, D] fp32
  float* part;           // [n_ctas][2][D] per-CTA column partials (sum, sum of squares); folded by fold_partials_kernel
  int fmt;               // 0 = fp16, 1 = bf16
};

template <int DP> struct GemmCfg {
  static constexpr int D = 64 * DP;
  static constexpr uint32_t A_PANEL = 128 * 128;        // 128 rows x 64 elements
  static constexpr uint32_t B_PANEL = D * 128;          // D rows x 64 elements
  static constexpr uint32_t STAGE = A_PANEL + B_PANEL;
  static constexpr int NSTAGE = 4;
  static constexpr uint32_t SMEM = NSTAGE * STAGE + 1024 + 256 + 4 * 2 * D * 4;
};

// warps: 0 TMA producer | 1 MMA issuer + TMEM owner | 2..5 epilogue (warp & 3 = TMEM lane quadrant)
template <int DP>
__global__ void __launch_bounds__(192, 1)
proj_gemm_kernel(const __grid_constant__ CUtensorMap tmap_h, const __grid_constant__ CUtensorMap tmap_w, GemmParams p) {
  using C = GemmCfg<DP>;
  constexpr int D = C::D, NSTAGE = C::NSTAGE;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t bars = base + NSTAGE * C::STAGE;
  auto bar_full = [&](int i) { return bars + 8u * i; };
  auto bar_empty = [&](int i) { return bars + 8u * (NSTAGE + i); };
  const uint32_t bar_done = bars + 8u * (2 * NSTAGE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + (bars - base) + 8u * (2 * NSTAGE + 1));
  float* wpart = reinterpret_cast<float*>(base_ptr + (bars - base) + 256);   // [4 warps][2][D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * 128;
  const int kt = p.K / 64;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(bar_full(i), 1); mbar_init(bar_empty(i), 1); }
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();
  pdl_launch();

  if (warp == 0) {
    if (elect_one()) { prefetch_tensormap(&tmap_h); prefetch_tensormap(&tmap_w); }
    for (int k = 0; k < kt; ++k) {
      const int s = k % NSTAGE;
      mbar_wait(bar_empty(s), ((uint32_t)(k / NSTAGE) & 1u) ^ 1u);
      if (elect_one()) {
        mbar_expect_tx(bar_full(s), C::STAGE);
        tma_load_2d(base + s * C::STAGE, &tmap_h, bar_full(s), k * 64, r0);               // rows past R are zero-filled
        tma_load_2d(base + s * C::STAGE + C::A_PANEL, &tmap_w, bar_full(s), k * 64, 0);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    const uint32_t idesc = idesc_16bit(128, D, p.fmt);
    constexpr uint32_t dhi = smem_desc_hi(1024);
    for (int k = 0; k < kt; ++k) {
      const int s = k % NSTAGE;
      mbar_wait(bar_full(s), (uint32_t)(k / NSTAGE) & 1u);
      tc_fence_after();
      const uint32_t a0 = smem_desc_lo(base + s * C::STAGE, 16);
      const uint32_t b0 = smem_desc_lo(base + s * C::STAGE + C::A_PANEL, 16);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          umma_ss(tmem, desc64(a0 + ((ks * 32) >> 4), dhi), desc64(b0 + ((ks * 32) >> 4), dhi), idesc, (k > 0 || ks > 0) ? 1u : 0u);
        umma_commit(bar_empty(s));
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(bar_done);
    __syncwarp();
  } else {
    // ---- epilogue: Y out, column sums of Y and Y^2 over this CTA's rows ----
    const int q = warp & 3;
    const int row = r0 + q * 32 + lane;
    const bool valid = row < p.R;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    float* wp = wpart + (warp - 2) * 2 * D;
    mbar_wait(bar_done, 0);
    tc_fence_after();
#pragma unroll
    for (int ch = 0; ch < 2 * DP; ++ch) {
      uint32_t v[32];
      tmem_ld_x32(tmem + lane_addr + ch * 32, v);
      tmem_ld_wait(v);
      float a[32], b[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float y = valid ? __uint_as_float(v[i]) : 0.f;       // rows past R contribute nothing to the statistics
        a[i] = y;
        b[i] = y * y;
      }
      if (valid) {
        float4* o = reinterpret_cast<float4*>(p.y + (size_t)row * D + ch * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = make_float4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]);
      }
      const float s1 = warp_transpose_reduce(a);
      const float s2 = warp_transpose_reduce(b);
      wp[ch * 32 + lane] = s1;
      wp[D + ch * 32 + lane] = s2;
    }
    named_bar_sync(1, 128);
    float* mine = p.part + (size_t)blockIdx.x * 2 * D;
    for (int i = threadIdx.x - 64; i < 2 * D; i += 128)
      mine[i] = (wpart[i] + wpart[2 * D + i]) + (wpart[4 * D + i] + wpart[6 * D + i]);      // fixed order over the 4 warps
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}

// ---- y^ = (Y - mean) rstd ; z = y^ / max(|y^|, eps) : one warp per row, D <= 256 ----
// totals = (sum, sumsq) over `count` rows (all ranks under SyncBatchNorm).  training: batch statistics + running-stat
// update (momentum, unbiased variance, as nn.BatchNorm1d); eval: the running statistics.
__global__ void __launch_bounds__(256)
proj_bn_l2_kernel(const float* __restrict__ y, int R, int D, const float* __restrict__ totals, float count, float bn_eps,
                  float l2_eps, int training, float momentum, float* __restrict__ running_mean,
                  float* __restrict__ running_var, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                  __nv_bfloat16* __restrict__ z, float* __restrict__ inv_norm) {
  __shared__ float s_mean[256], s_rstd[256];
  pdl_wait();
  pdl_launch();
  for (int c = threadIdx.x; c < D; c += blockDim.x) {
    float m, var;
    if (training) {
      m = totals[c] / count;
      var = fmaxf(totals[D + c] / count - m * m, 0.f);            // biased variance normalises the batch
      if (blockIdx.x == 0 && running_mean != nullptr) {
        const float unbiased = count > 1.f ? var * count / (count - 1.f) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * m;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
      }
    } else {
      m = running_mean[c];
      var = running_var[c];
    }
    const float r = rsqrtf(var + bn_eps);
    s_mean[c] = m;
    s_rstd[c] = r;
    if (blockIdx.x == 0) { mean_out[c] = m; rstd_out[c] = r; }
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= R) return;
  const bool have = lane < D / 8;
  float v[8];
  float ss = 0.f;
  if (have) {
    float a[4], b[4];
    VecIO<float>::load(y + (size_t)row * D + lane * 8, a);
    VecIO<float>::load(y + (size_t)row * D + lane * 8 + 4, b);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[i] = (a[i] - s_mean[lane * 8 + i]) * s_rstd[lane * 8 + i];
      v[4 + i] = (b[i] - s_mean[lane * 8 + 4 + i]) * s_rstd[lane * 8 + 4 + i];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) ss = fmaf(v[i], v[i], ss);
  }
  ss = warp_sum(ss);
  const float inv = 1.0f / fmaxf(sqrtf(ss), l2_eps);
  if (have) {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] *= inv;
    VecIO<__nv_bfloat16>::store(z + (size_t)row * D + lane * 8, v);
  }
  if (lane == 0) inv_norm[row] = inv;
}

// ---- totals[i] = sum over CTAs of part[cta][i], fixed order: one warp per value, lane l takes CTAs l, l + 32, ... and the
// 32 lane sums are folded by a shuffle tree (a serial fold by one thread costs an L2 round trip per CTA: measured 160 us
// for 512 CTAs) ----
__global__ void __launch_bounds__(256)
fold_partials_kernel(const float* __restrict__ part, int n_ctas, int n_vals, float* __restrict__ totals) {
  pdl_wait();
  pdl_launch();
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (i >= n_vals) return;
  double s = 0.0;
  for (int b = lane; b < n_ctas; b += 32) s += (double)__ldg(part + (size_t)b * n_vals + i);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) totals[i] = (float)s;
}

// ---- backward pass 1: dy^ = inv (dz - z <z, dz>) per row (rows on the eps clamp: inv dz); column partial sums of dy^ and
// dy^ * y^ (BatchNorm backward); dy^ stored fp32.  One warp per row, CTAs loop over 8-row groups and keep their column
// sums in registers; per-CTA partials are folded by fold_partials_kernel. ----
__global__ void __launch_bounds__(256)
proj_bwd1_kernel(const float* __restrict__ dz_partials, int n_partials, int64_t partial_stride, const __nv_bfloat16* __restrict__ z,
                 const float* __restrict__ inv_norm, float inv_eps, const float* __restrict__ y, const float* __restrict__ mean,
                 const float* __restrict__ rstd, int R, int D, float* __restrict__ dyhat, float* __restrict__ part) {
  __shared__ float s_part[8][2][256];
  pdl_wait();
  pdl_launch();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const bool col_ok = lane < D / 8;
  float mu[8], rs[8], c1[8], c2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    mu[i] = col_ok ? mean[lane * 8 + i] : 0.f;
    rs[i] = col_ok ? rstd[lane * 8 + i] : 0.f;
    c1[i] = 0.f; c2[i] = 0.f;
  }
  for (int row = blockIdx.x * 8 + w; row < R; row += gridDim.x * 8) {
    float g[8], zv[8], yh[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { g[i] = 0.f; zv[i] = 0.f; yh[i] = 0.f; }
    if (col_ok) {
      for (int k = 0; k < n_partials; ++k) {
        float a[4], b[4];
        VecIO<float>::load(dz_partials + k * partial_stride + (size_t)row * D + lane * 8, a);
        VecIO<float>::load(dz_partials + k * partial_stride + (size_t)row * D + lane * 8 + 4, b);
#pragma unroll
        for (int i = 0; i < 4; ++i) { g[i] += a[i]; g[4 + i] += b[i]; }
      }
      VecIO<__nv_bfloat16>::load(z + (size_t)row * D + lane * 8, zv);
      float a[4], b[4];
      VecIO<float>::load(y + (size_t)row * D + lane * 8, a);
      VecIO<float>::load(y + (size_t)row * D + lane * 8 + 4, b);
#pragma unroll
      for (int i = 0; i < 4; ++i) { yh[i] = (a[i] - mu[i]) * rs[i]; yh[4 + i] = (b[i] - mu[4 + i]) * rs[4 + i]; }
    }
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) dot = fmaf(g[i], zv[i], dot);
    dot = warp_sum(dot);
    const float inv = inv_norm[row];
    if (inv >= inv_eps) dot = 0.f;                              // |y^| <= eps: F.normalize divides by the constant eps
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      g[i] = (g[i] - zv[i] * dot) * inv;
      c1[i] += g[i];
      c2[i] = fmaf(g[i], yh[i], c2[i]);
    }
    if (col_ok) {
      VecIO<float>::store(dyhat + (size_t)row * D + lane * 8, reinterpret_cast<float(&)[4]>(g[0]));
      VecIO<float>::store(dyhat + (size_t)row * D + lane * 8 + 4, reinterpret_cast<float(&)[4]>(g[4]));
    }
  }
  if (col_ok) {
#pragma unroll
    for (int i = 0; i < 8; ++i) { s_part[w][0][lane * 8 + i] = c1[i]; s_part[w][1][lane * 8 + i] = c2[i]; }
  }
  __syncthreads();
  float* mine = part + (size_t)blockIdx.x * 2 * D;
  for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) {
    const int which = i / D, c = i - which * D;
    float s2 = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s2 += s_part[k][which][c];
    mine[i] = s2;
  }
}

// ---- backward pass 2: dY = rstd (dy^ - s1 / count - y^ s2 / count)   (eval mode: dY = rstd dy^) ----
template <typename TOut>
__global__ void __launch_bounds__(256)
proj_bwd2_kernel(const float* __restrict__ dyhat, const float* __restrict__ y, const float* __restrict__ mean,
                 const float* __restrict__ rstd, const float* __restrict__ totals, float count, int training, int R, int D,
                 TOut* __restrict__ dy) {
  pdl_wait();
  pdl_launch();
  const int64_t total = (int64_t)R * D / 8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c0 = (int)((i * 8) % D);
    float g[8], yy[8];
    VecIO<float>::load(dyhat + i * 8, reinterpret_cast<float(&)[4]>(g[0]));
    VecIO<float>::load(dyhat + i * 8 + 4, reinterpret_cast<float(&)[4]>(g[4]));
    VecIO<float>::load(y + i * 8, reinterpret_cast<float(&)[4]>(yy[0]));
    VecIO<float>::load(y + i * 8 + 4, reinterpret_cast<float(&)[4]>(yy[4]));
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float r = rstd[c0 + k];
      if (training) {
        const float yh = (yy[k] - mean[c0 + k]) * r;
        o[k] = r * (g[k] - totals[c0 + k] / count - yh * totals[D + c0 + k] / count);
      } else {
        o[k] = r * g[k];
      }
    }
    if constexpr (sizeof(TOut) == 4) {
      VecIO<float>::store(reinterpret_cast<float*>(dy) + i * 8, reinterpret_cast<float(&)[4]>(o[0]));
      VecIO<float>::store(reinterpret_cast<float*>(dy) + i * 8 + 4, reinterpret_cast<float(&)[4]>(o[4]));
    } else {
      VecIO<TOut>::store(dy + i * 8, o);
    }
  }
}

}  // namespace
}  // namespace sm3

using namespace sm3;

extern "C" int sm3_proj_tail_supported(int K, int D, int dtype) {
  static int sm100 = -1;                                  // device query once: this is asked on every forward
  if (sm100 < 0) sm100 = sm3_device_supported() == 1 ? 1 : 0;
  return ((dtype == SM3_F16 || dtype == SM3_BF16) && K % 64 == 0 && K >= 64 && D % 64 == 0 && D >= 64 && D <= 256 &&
          sm100 == 1) ? 1 : 0;
}

static int64_t proj_bwd1_ctas(int64_t R) {
  const int64_t want = (R + 7) / 8, cap = (int64_t)num_sms() * 4;
  return want < cap ? want : cap;
}
extern "C" size_t sm3_proj_tail_workspace_bytes(int64_t R, int D) {
  const int64_t a = (R + 127) / 128, b = proj_bwd1_ctas(R);            // per-CTA column partials of the GEMM / of bwd1
  return (size_t)(a > b ? a : b) * 2 * D * sizeof(float) + 256;
}

// Y [R, D] fp32 = H [R, K] W[D, K]^T (16-bit operands, fp32 accumulation on the tensor cores) and totals[2 D] =
// (sum_r Y, sum_r Y^2).  workspace from sm3_proj_tail_workspace_bytes.
extern "C" int sm3_proj_tail_gemm(const void* h, const void* w, int64_t R, int K, int D, int dtype, float* y, float* totals,
                                  void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SM3_REQUIRE(h && w && y && totals && workspace, SM3_ERR_SHAPE, "proj_tail_gemm: null pointer");
  SM3_REQUIRE(R >= 1 && R < ((int64_t)1 << 31), SM3_ERR_SHAPE, "proj_tail_gemm: bad row count");
  SM3_REQUIRE(sm3_proj_tail_supported(K, D, dtype), SM3_ERR_DTYPE,
              "proj_tail_gemm: needs fp16 / bf16 operands, K %% 64 == 0, D in {64,128,192,256} (got K=%d D=%d dtype=%d)", K, D, dtype);
  SM3_REQUIRE(aligned16(h) && aligned16(w) && aligned16(y), SM3_ERR_SHAPE, "proj_tail_gemm: unaligned pointer");
  SM3_REQUIRE(workspace_bytes >= sm3_proj_tail_workspace_bytes(R, D), SM3_ERR_WORKSPACE, "proj_tail_gemm: workspace too small");
  CUtensorMap th, tw;
  int rc = make_tmap_16bit(&th, h, (uint64_t)R, (uint64_t)K, 128, dtype == SM3_F16);
  if (rc) return rc;
  if ((rc = make_tmap_16bit(&tw, w, (uint64_t)D, (uint64_t)K, (uint32_t)D, dtype == SM3_F16))) return rc;
  const int ctas = (int)((R + 127) / 128);
  GemmParams p{(int)R, K, D, y, (float*)workspace, dtype == SM3_F16 ? 0 : 1};
#define SM3_PG(DP)                                                                                                     \
  do {                                                                                                                 \
    SM3_CHECK_CUDA(cudaFuncSetAttribute(proj_gemm_kernel<DP>, cudaFuncAttributeMaxDynamicSharedMemorySize,             \
                                        (int)GemmCfg<DP>::SMEM));                                                      \
    SM3_CHECK_CUDA(launch_k(proj_gemm_kernel<DP>, dim3(ctas), dim3(192), GemmCfg<DP>::SMEM, st, th, tw, p));           \
  } while (0)
  switch (D / 64) {
    case 1: SM3_PG(1); break;
    case 2: SM3_PG(2); break;
    case 3: SM3_PG(3); break;
    default: SM3_PG(4); break;
  }
#undef SM3_PG
  SM3_CHECK_CUDA(launch_k(fold_partials_kernel, dim3((2 * D + 7) / 8), dim3(256), 0, st, (const float*)workspace, ctas, 2 * D, totals));
  return SM3_OK;
}

// z (bf16) = normalize(batchnorm(Y)), inv_norm; mean / rstd [D] saved for the backward; running statistics updated in
// training mode (may be NULL: track_running_stats=False).  `count` = rows behind `totals` (all ranks under SyncBN).
extern "C" int sm3_proj_tail_bn_l2(const float* y, int64_t R, int D, const float* totals, float count, float bn_eps,
                                   float l2_eps, int training, float momentum, float* running_mean, float* running_var,
                                   float* mean_out, float* rstd_out, void* z_bf16, float* inv_norm, void* stream) {
  SM3_REQUIRE(y && z_bf16 && inv_norm && mean_out && rstd_out, SM3_ERR_SHAPE, "proj_tail_bn_l2: null pointer");
  SM3_REQUIRE(training ? totals != nullptr : (running_mean != nullptr && running_var != nullptr), SM3_ERR_SHAPE,
              "proj_tail_bn_l2: statistics missing");
  SM3_REQUIRE(R >= 1 && D % 8 == 0 && D >= 8 && D <= 256 && count >= 1.f, SM3_ERR_SHAPE, "proj_tail_bn_l2: bad shape");
  SM3_CHECK_CUDA(launch_k(proj_bn_l2_kernel, dim3((unsigned)((R + 7) / 8)), dim3(256), 0, (cudaStream_t)stream, y, (int)R, D,
                          totals, count, bn_eps, l2_eps, training, momentum, running_mean, running_var, mean_out, rstd_out,
                          (__nv_bfloat16*)z_bf16, inv_norm));
  return SM3_OK;
}

// backward pass 1 (see proj_bwd1_kernel): dyhat [R, D] fp32 and totals2[2 D] = (sum_r dy^, sum_r dy^ y^)
extern "C" int sm3_proj_tail_bwd1(const float* dz_partials, int n_partials, int64_t partial_stride, const void* z_bf16,
                                  const float* inv_norm, float l2_eps, const float* y, const float* mean, const float* rstd,
                                  int64_t R, int D, float* dyhat, float* totals2, void* workspace, size_t workspace_bytes,
                                  void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SM3_REQUIRE(dz_partials && z_bf16 && inv_norm && y && mean && rstd && dyhat && totals2 && workspace, SM3_ERR_SHAPE,
              "proj_tail_bwd1: null pointer");
  SM3_REQUIRE(R >= 1 && D % 8 == 0 && D <= 256 && n_partials >= 1, SM3_ERR_SHAPE, "proj_tail_bwd1: bad shape");
  SM3_REQUIRE(workspace_bytes >= sm3_proj_tail_workspace_bytes(R, D), SM3_ERR_WORKSPACE, "proj_tail_bwd1: workspace too small");
  const int ctas = (int)proj_bwd1_ctas(R);
  SM3_CHECK_CUDA(launch_k(proj_bwd1_kernel, dim3((unsigned)ctas), dim3(256), 0, st, dz_partials, n_partials, partial_stride,
                          (const __nv_bfloat16*)z_bf16, inv_norm, 1.0f / l2_eps, y, mean, rstd, (int)R, D, dyhat,
                          (float*)workspace));
  SM3_CHECK_CUDA(launch_k(fold_partials_kernel, dim3((2 * D + 7) / 8), dim3(256), 0, st, (const float*)workspace, ctas, 2 * D,
                          totals2));
  return SM3_OK;
}

// backward pass 2: dY (dtype of H) from dyhat and the (all-reduced) column sums
extern "C" int sm3_proj_tail_bwd2(const float* dyhat, const float* y, const float* mean, const float* rstd, const float* totals2,
                                  float count, int training, int64_t R, int D, void* dy, int dy_dtype, void* stream) {
  SM3_REQUIRE(dyhat && y && mean && rstd && dy && (totals2 || !training), SM3_ERR_SHAPE, "proj_tail_bwd2: null pointer");
  SM3_REQUIRE(R >= 1 && D % 8 == 0 && dtype_ok(dy_dtype) && count >= 1.f, SM3_ERR_SHAPE, "proj_tail_bwd2: bad shape");
  int64_t blocks = ((R * D / 8) + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  SM3_DISPATCH_DTYPE(dy_dtype, TOut, {
    SM3_CHECK_CUDA(launch_k(proj_bwd2_kernel<TOut>, dim3((unsigned)blocks), dim3(256), 0, (cudaStream_t)stream, dyhat, y, mean,
                            rstd, totals2, count, training, (int)R, D, (TOut*)dy));
  });
  return SM3_OK;
}
