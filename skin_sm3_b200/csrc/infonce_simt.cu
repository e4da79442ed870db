// K2/K3, exact-fp32 variant: fused similarity + InfoNCE statistics (forward) and the row-local symmetric
// backward, on the FP32 FMA pipes.  Used when fp32 parity with the reference is asked for (loss 1e-5 /
// grads 1e-4 relative), for embedding widths the tensor-core path does not take, and as the on-GPU
// cross-check of the tcgen05 kernels.  Same maths as infonce_tc.cu; see sm3_b200.h for the contract.
//
// Reference being replaced: src/models/simclr.py:296-320 (+ :64-88, :140-164) and the CE at
// tools/backbone_train.py:531 / its autograd backward.  Nothing of size [M, M] is ever written.
//
// Tiling: one CTA = 64 rows x a contiguous range of 64-column tiles; both 64 x D operand tiles live in
// shared memory (row stride D+1 floats: conflict-free), each thread owns a 4 x 4 micro-tile of S.
// Backward adds a second phase per tile: H (64 x 64, shared memory) times the column tile -> 64 x D
// accumulators in registers.
#include "common.cuh"

namespace sm3 {
namespace {

constexpr int kTile = 64;
constexpr int kThreads = 256;
constexpr float kLog2e = 1.4426950408889634f;

struct SimtParams {
  const void* z_rows;
  const void* z_cols;
  int n_local, pair_offset, n_global, D;
  int m_rows, m_cols;
  float inv_T, c2;               // c2 = inv_T * log2(e)
  int tiles_per_split, col_tiles;
  // forward
  float* pos;
  float* partial;                // [splits][m_rows]
  // backward
  const float *gpos_r, *glse_r, *nsum_r, *gpos_c, *glse_c, *nsum_c;
  int cstride;                   // element stride of the *_c arrays
  float* dz_partial;             // [splits][m_rows][D]
};

__device__ __forceinline__ float safe_coef(float g, float s) { return s > 0.f ? g / s : 0.f; }

template <typename T, int DPAD, bool BWD>
__global__ void __launch_bounds__(kThreads)
infonce_simt_kernel(SimtParams p) {
  extern __shared__ float smem[];
  constexpr int LD = DPAD + 1;
  float* zr = smem;                       // [64][LD]
  float* zc = zr + kTile * LD;            // [64][LD]
  float* hs = zc + kTile * LD;            // [64][65]   (backward only)
  float* ac = hs + (BWD ? kTile * (kTile + 1) : 0);   // [64] a_j        (backward only)
  float* gc = ac + (BWD ? kTile : 0);                 // [64] g_pos_j    (backward only)

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int r0 = blockIdx.x * kTile;
  const int split = blockIdx.y;
  const int D = p.D;
  const T* zrows = reinterpret_cast<const T*>(p.z_rows);
  const T* zcols = reinterpret_cast<const T*>(p.z_cols);

  // ---- stage the 64 x D row tile once (zero padded) ----
  for (int idx = tid; idx < kTile * DPAD; idx += kThreads) {
    const int r = idx / DPAD, k = idx - r * DPAD;
    const int l = r0 + r;
    zr[r * LD + k] = (l < p.m_rows && k < D) ? to_f32(zrows[(int64_t)l * D + k]) : 0.f;
  }

  int grow[4], gposi[4];
  float a_row[4], gp_row[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int l = r0 + ty + 16 * i;
    const bool ok = l < p.m_rows;
    grow[i] = ok ? global_row(l, p.n_local, p.pair_offset, p.n_global) : -1;
    gposi[i] = ok ? positive_of(grow[i], p.n_global) : -1;
    if constexpr (BWD) {
      a_row[i] = ok ? safe_coef(p.glse_r[l], p.nsum_r[l]) : 0.f;
      gp_row[i] = ok ? p.gpos_r[l] : 0.f;
    }
  }

  float rowsum[4] = {0.f, 0.f, 0.f, 0.f};
  constexpr int NC = DPAD / 16;
  float acc2[BWD ? 4 : 1][BWD ? NC : 1];
  if constexpr (BWD) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int c = 0; c < NC; ++c) acc2[i][c] = 0.f;
  }

  const int t_begin = split * p.tiles_per_split;
  const int t_end = min(p.col_tiles, t_begin + p.tiles_per_split);
  for (int t = t_begin; t < t_end; ++t) {
    const int c0 = t * kTile;
    __syncthreads();   // previous tile fully consumed (also orders the zr fill on the first pass)
    for (int idx = tid; idx < kTile * DPAD; idx += kThreads) {
      const int r = idx / DPAD, k = idx - r * DPAD;
      const int j = c0 + r;
      zc[r * LD + k] = (j < p.m_cols && k < D) ? to_f32(zcols[(int64_t)j * D + k]) : 0.f;
    }
    if constexpr (BWD) {
      if (tid < kTile) {
        const int j = c0 + tid;
        ac[tid] = j < p.m_cols ? safe_coef(p.glse_c[(size_t)j * p.cstride], p.nsum_c[(size_t)j * p.cstride]) : 0.f;
        gc[tid] = j < p.m_cols ? p.gpos_c[(size_t)j * p.cstride] : 0.f;
      }
    }
    __syncthreads();

    // ---- phase 1: 4 x 4 micro-tile of S = Zr Zc^T ----
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
#pragma unroll 8
    for (int k = 0; k < DPAD; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = zr[(ty + 16 * i) * LD + k];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = zc[(tx + 16 * j) * LD + k];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }

    // ---- epilogue: mask diagonal / positive, shifted exponential ----
#pragma unroll
    for (int i = 0; i < 4; ++i) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int col = c0 + tx + 16 * j;
        const float s = acc[i][j];
        float h = 0.f;
        if (grow[i] >= 0 && col < p.m_cols && col != grow[i]) {
          if (col == gposi[i]) {
            if constexpr (BWD) h = gp_row[i] + gc[tx + 16 * j];
            else p.pos[r0 + ty + 16 * i] = s * p.inv_T;
          } else {
            const float e = exp2f(fmaf(s, p.c2, -p.c2));
            if constexpr (BWD) h = e * (a_row[i] + ac[tx + 16 * j]);
            else rowsum[i] += e;
          }
        }
        if constexpr (BWD) hs[(ty + 16 * i) * (kTile + 1) + tx + 16 * j] = h;
      }
    }

    if constexpr (BWD) {
      __syncthreads();
      // ---- phase 2: dZ[64 x D] += H[64 x 64] * Zc[64 x D] ----
#pragma unroll 4
      for (int j = 0; j < kTile; ++j) {
        float h[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = hs[(ty + 16 * i) * (kTile + 1) + j];
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const float zv = zc[j * LD + tx + 16 * c];
#pragma unroll
          for (int i = 0; i < 4; ++i) acc2[i][c] = fmaf(h[i], zv, acc2[i][c]);
        }
      }
    }
  }

  if constexpr (!BWD) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float v = rowsum[i];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      v += __shfl_xor_sync(0xffffffffu, v, 4);
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      const int l = r0 + ty + 16 * i;
      if (tx == 0 && l < p.m_rows) p.partial[(int64_t)split * p.m_rows + l] = v;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int l = r0 + ty + 16 * i;
      if (l < p.m_rows) {
        float* dst = p.dz_partial + ((int64_t)split * p.m_rows + l) * D;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const int d = tx + 16 * c;
          if (d < D) dst[d] = acc2[i][c] * p.inv_T;
        }
      }
    }
  }
}

int pick_splits(int row_tiles, int col_tiles) {
  int want = (2 * num_sms() + row_tiles - 1) / row_tiles;
  if (want < 1) want = 1;
  if (want > 16) want = 16;
  if (want > col_tiles) want = col_tiles;
  const int tps = (col_tiles + want - 1) / want;
  return (col_tiles + tps - 1) / tps;   // no empty split
}

size_t smem_bytes(int dpad, bool bwd) {
  size_t f = 2 * (size_t)kTile * (dpad + 1);
  if (bwd) f += (size_t)kTile * (kTile + 1) + 2 * kTile;
  return f * sizeof(float);
}

template <typename T, int DPAD, bool BWD>
int launch_one(const SimtParams& p, int row_tiles, int splits, cudaStream_t st) {
  const size_t smem = smem_bytes(DPAD, BWD);
  SM3_CHECK_CUDA(cudaFuncSetAttribute(infonce_simt_kernel<T, DPAD, BWD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
  infonce_simt_kernel<T, DPAD, BWD><<<dim3(row_tiles, splits), kThreads, smem, st>>>(p);
  SM3_CHECK_CUDA(cudaGetLastError());
  return SM3_OK;
}

template <bool BWD>
int launch(const SimtParams& p, int dtype, int row_tiles, int splits, cudaStream_t st) {
  const int D = p.D;
  int rc = SM3_OK;
  SM3_DISPATCH_DTYPE(dtype, T, {
    if (D <= 64) rc = launch_one<T, 64, BWD>(p, row_tiles, splits, st);
    else if (D <= 128) rc = launch_one<T, 128, BWD>(p, row_tiles, splits, st);
    else rc = launch_one<T, 256, BWD>(p, row_tiles, splits, st);
  });
  return rc;
}

int fill(const InfoNceProblem& pb, SimtParams& p, int& row_tiles, int& splits) {
  SM3_REQUIRE(pb.D >= 1 && pb.D <= 256, SM3_ERR_DTYPE, "infonce(simt): D=%d not in [1,256]", pb.D);
  SM3_REQUIRE(pb.n_local >= 1 && pb.n_global >= pb.n_local && pb.pair_offset >= 0 &&
                  pb.pair_offset + pb.n_local <= pb.n_global,
              SM3_ERR_SHAPE, "infonce: bad row block (n_local=%d offset=%d n_global=%d)", pb.n_local, pb.pair_offset,
              pb.n_global);
  SM3_REQUIRE(pb.n_global <= (1 << 29), SM3_ERR_SHAPE, "infonce: n_global too large");
  p.z_rows = pb.z_rows; p.z_cols = pb.z_cols;
  p.n_local = pb.n_local; p.pair_offset = pb.pair_offset; p.n_global = pb.n_global; p.D = pb.D;
  p.m_rows = 2 * pb.n_local; p.m_cols = 2 * pb.n_global;
  p.inv_T = pb.inv_T; p.c2 = pb.inv_T * kLog2e;
  row_tiles = (p.m_rows + kTile - 1) / kTile;
  p.col_tiles = (p.m_cols + kTile - 1) / kTile;
  splits = pick_splits(row_tiles, p.col_tiles);
  p.tiles_per_split = (p.col_tiles + splits - 1) / splits;
  return SM3_OK;
}

}  // namespace

size_t infonce_simt_workspace(const InfoNceProblem& pb, int backward) {
  const int m_rows = 2 * pb.n_local, m_cols = 2 * pb.n_global;
  const int row_tiles = (m_rows + kTile - 1) / kTile, col_tiles = (m_cols + kTile - 1) / kTile;
  const int splits = pick_splits(row_tiles, col_tiles);
  const size_t per = backward ? (size_t)m_rows * pb.D : (size_t)m_rows;
  return (size_t)splits * per * sizeof(float) + 256;
}

int infonce_simt_fwd(const InfoNceProblem& pb, float* pos, float* lse_neg, float* neg_sum, void* ws, size_t ws_bytes,
                     cudaStream_t st) {
  SimtParams p{};
  int row_tiles, splits;
  int rc = fill(pb, p, row_tiles, splits);
  if (rc) return rc;
  SM3_REQUIRE(ws_bytes >= infonce_simt_workspace(pb, 0), SM3_ERR_WORKSPACE, "infonce(simt) fwd: workspace too small");
  p.pos = pos;
  p.partial = (float*)ws;
  rc = launch<false>(p, pb.dtype, row_tiles, splits, st);
  if (rc) return rc;
  return infonce_finalize_launch(p.partial, splits, p.m_rows, pb.inv_T, neg_sum, lse_neg, st);
}

int infonce_simt_bwd(const InfoNceProblem& pb, const float* gpos_r, const float* glse_r, const float* nsum_r,
                     const float* gpos_c, const float* glse_c, const float* nsum_c, void* ws, size_t ws_bytes,
                     cudaStream_t st) {
  SimtParams p{};
  int row_tiles, splits;
  int rc = fill(pb, p, row_tiles, splits);
  if (rc) return rc;
  SM3_REQUIRE(ws_bytes >= infonce_simt_workspace(pb, 1), SM3_ERR_WORKSPACE, "infonce(simt) bwd: workspace too small");
  p.gpos_r = gpos_r; p.glse_r = glse_r; p.nsum_r = nsum_r;
  p.gpos_c = gpos_c; p.glse_c = glse_c; p.nsum_c = nsum_c; p.cstride = pb.col_stride;
  p.dz_partial = (float*)ws;
  rc = launch<true>(p, pb.dtype, row_tiles, splits, st);
  if (rc) return rc;
  return splits;
}

}  // namespace sm3
