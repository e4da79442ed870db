// K2/K3 on the 5th-generation tensor cores: flash-style fused similarity + InfoNCE for bf16 embeddings.
//
//   forward  : S-tile = Zr (128 x D, resident in TMEM) * Zc^T (128 x D tile, TMA -> smem, SWIZZLE_128B)
//              accumulated in TMEM, read back with tcgen05.ld by 8 softmax warps (one thread = one row, no
//              shuffles), exp2 with the constant shift 1/T (|s| <= 1 for unit rows), running row sums.
//   backward : per 128 x 64 tile   S = Zr Zc^T (tcgen05, A from TMEM)  ->  H = exp(.)(a_i + a_j) as bf16
//              written back into the S columns of TMEM  ->  dZ (128 x D, TMEM) += H * Zc, the SAME smem tile
//              re-read as an MN-major B operand.  Row-local thanks to the symmetry of S (see sm3_b200.h).
//   Nothing of size [M, M] is ever written; HBM traffic is ~ M*D*2 bytes per row-block pass (L2 resident).
//
// Kernels in this file:
//   infonce_tc_fwd_kernel        forward, 128-row CTAs (small problems)
//   infonce_tc_fwd2_kernel       forward, 256-row CTAs: every B tile feeds two S-MMAs (multi-rank row blocks; optional
//                                owner-ordered tiles + in-kernel row push)
//   infonce_tc_fwdsym_kernel     forward on ONE rank: S is symmetric, only the upper-triangular tiles are computed and a
//                                tile's column sums stand in for the transposed tile (flat persistent work list)
//   infonce_tc_fwdsym_mr_kernel  the same ACROSS ranks: W/2 of the W column blocks per rank (circulant assignment),
//                                column sums of foreign blocks shipped to their owners (exchange mode 4)
//   infonce_tc_bwd_kernel        backward, 64-column tiles (D > 128: TMEM is full)
//   infonce_tc_bwd2_kernel       backward, 128-column tiles / tile-alternating groups (D <= 128)
//
// Replaces: matmul + mask/gather/cat + /T of src/models/simclr.py:296-320 (= :64-88, :140-164), the CE of
// tools/backbone_train.py:531 and the autograd backward of both.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = tcgen05.mma issuer + TMEM owner,
// warps 2..9 = softmax/epilogue (warp_id % 4 selects the TMEM lane quadrant, the pair of warps sharing a
// quadrant split the tile's columns).  Pipelines: smem full/empty (TMA <-> MMA), TMEM S full/empty
// (MMA <-> softmax), and in the backward H-ready (softmax -> MMA).
#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "ptx_sm100.cuh"

namespace sm3 {
namespace {

using namespace ptx;

constexpr int kBM = 128;
constexpr float kLog2eTC = 1.4426950408889634f;

// Optional pipeline trace (build with -DSM3_TRACE: `make TRACE=1` -> lib/libsm3_b200_trace.so).  CTA (0,0) stamps
// clock64() at the hand-off points of the first tiles; tools/trace_tc.py prints the timeline.  Compiled out otherwise.
#ifdef SM3_TRACE
constexpr int kTraceTiles = 512, kTraceKinds = 12;
__device__ long long g_trace[kTraceTiles * kTraceKinds];
#define SM3_TR(kind, it)                                                                 \
  do {                                                                                   \
    if (blockIdx.x == 0 && blockIdx.y == 0 && (it) < kTraceTiles) g_trace[(it) * kTraceKinds + (kind)] = clock64(); \
  } while (0)
#else
#define SM3_TR(kind, it) do { } while (0)
#endif

struct TcParams {
  int n_local, pair_offset, n_global, D;
  int m_rows, m_cols;
  float inv_T, c2;
  int tiles_per_split, col_tiles;   // col_tiles counts the tiles actually visited (local tiles excluded in skip mode)
  int skip_a, skip_b, skip_len;     // skip mode: original tile indices of the two local ranges and their length
  const __nv_bfloat16* z_rows;
  float* pos;
  float* partial;      // fwd: [splits][m_rows]
  const float *gpos_r, *glse_r, *nsum_r, *gpos_c;
  int cstride;         // element stride of gpos_c (1, or 4 for packed peer-exchanged statistics)
  const float* acol;   // bwd: a_j = g_lse_j / neg_sum_j, zero padded to a multiple of 64
  float* dz_partial;   // bwd: [splits][m_rows][D]
  // fused exchange: flags to wait for before the first column tile owned by another rank (nullptr = no waiting), and
  // the two local column-tile ranges [loc_a, loc_a + loc_len), [loc_b, loc_b + loc_len) in this kernel's tile units
  const unsigned* wait_flags;
  int wait_world, wait_channel;
  unsigned wait_epoch;
  unsigned long long wait_timeout_ns;
  int loc_a, loc_b, loc_len;
  // owner-ordered tiles + in-kernel row push (multi-rank forward, see infonce_tc_fwd2_kernel): tiles are visited owner by
  // owner -- this rank's own columns, then rank-1's, rank-2's, ... -- and split s takes every own_S-th tile of an
  // owner's two ranges (own_L tiles each); the kernel's two extra warps store this rank's rows into the peers' column
  // buffers destination by destination in the order rank+1, rank+2, ... and release a per-destination flag.
  int own_order, own_L, own_S;
  const uint4* push_src;         // this rank's normalised rows [m_rows, D] (nullptr = no pushing)
  int push_vpr, push_ctas;       // 16-byte vectors per row; how many CTAs (the first wave) share the copy
  PeerPtrs push_dst, push_flags;
  unsigned* push_counter;        // [world] zero-initialised tickets, reset by the last CTA
  // symmetric forward (single rank): column tiles per side, tiles per CTA, row-sum slabs, flat work-list length
  int sym_T, sym_tpc, sym_maxseg;
  long sym_W;
  MrPlan mr;            // cross-rank symmetric forward (infonce_tc_fwdsym_mr_kernel)
};

// Per-CTA starting rotation of the column-tile order.  Default: pseudo-random (decorrelates the CTAs of a wave, see the
// forward kernel).  Fused exchange: start inside this rank's own column range when the CTA's split touches it, so the
// tiles that need no remote data are visited while the peers' rows are still in flight.
__device__ __forceinline__ int tile_rotation(const TcParams& p, int t_begin, int n_tiles) {
  if (n_tiles <= 0) return 0;
  const unsigned h = blockIdx.x * 37u + blockIdx.y * 11u;
  if (p.wait_flags != nullptr) {
    const int t_end = t_begin + n_tiles;
    const int s0 = max(t_begin, p.loc_a), e0 = min(t_end, p.loc_a + p.loc_len);
    if (s0 < e0) return s0 - t_begin + (int)(h % (unsigned)(e0 - s0));
    const int s1 = max(t_begin, p.loc_b), e1 = min(t_end, p.loc_b + p.loc_len);
    if (s1 < e1) return s1 - t_begin + (int)(h % (unsigned)(e1 - s1));
  }
  return (int)(h % (unsigned)n_tiles);
}
__device__ __forceinline__ bool tile_is_local(const TcParams& p, int t) {
  return (unsigned)(t - p.loc_a) < (unsigned)p.loc_len || (unsigned)(t - p.loc_b) < (unsigned)p.loc_len;
}
// owner-ordered tile of visiting position `it` for split `split` (see TcParams::own_order); *owner = rank that owns it
__device__ __forceinline__ int owner_tile(const TcParams& p, int it, int split, int rot, int* owner) {
  const int per = p.own_L / p.own_S;           // tiles of one owner range that one split visits
  const int j = it / (2 * per);                // owner's position in the visiting order
  const int r = it - j * 2 * per;
  const int half = r >= per ? 1 : 0;
  int q = r - half * per + rot;                // per-CTA rotation inside the range (decorrelates the CTAs' L2 accesses)
  if (q >= per) q -= per;
  int o = p.wait_world > 0 ? (p.pair_offset / p.n_local) - j : 0;
  if (o < 0) o += p.wait_world;
  *owner = o;
  return half * (p.wait_world * p.own_L) + o * p.own_L + split + q * p.own_S;
}
// one source rank's flag (channel * 16 + src) instead of all of them
__device__ __forceinline__ void peer_flag_wait_one(const unsigned* flags, int channel, int src, unsigned epoch,
                                                   unsigned long long timeout_ns) {
  peer_flags_wait_all(flags + src, 1, channel, epoch, timeout_ns);
}
// generic-proxy acquire of the peers' flags -> async-proxy (TMA) reads of the rows they published
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// TMA descriptor (driver entry point fetched through the runtime: no link-time libcuda dependency)
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  });
  return fn;
}

// row-major bf16 matrix [rows, cols]; box = 64 columns (128 B, one swizzle row) x box_rows rows
int make_tmap_bf16_uncached(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows);
struct TmapKey { const void* base; uint64_t rows, cols; uint32_t box_rows; };
int make_tmap_bf16(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  // the encoded descriptor depends only on (pointer, shape, box): steps that reuse their scratch hit this cache
  // (cuTensorMapEncodeTiled is ~3 us of host time, twice per step)
  static thread_local TmapKey keys[8];
  static thread_local CUtensorMap vals[8];
  static thread_local int next = 0;
  for (int i = 0; i < 8; ++i)
    if (keys[i].base == base && keys[i].rows == rows && keys[i].cols == cols && keys[i].box_rows == box_rows && base != nullptr) {
      *map = vals[i];
      return SM3_OK;
    }
  const int rc = make_tmap_bf16_uncached(map, base, rows, cols, box_rows);
  if (rc == SM3_OK) { keys[next] = TmapKey{base, rows, cols, box_rows}; vals[next] = *map; next = (next + 1) & 7; }
  return rc;
}
int make_tmap_bf16_uncached(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  SM3_REQUIRE(fn != nullptr, SM3_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {cols * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SM3_REQUIRE(r == CUDA_SUCCESS, SM3_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu box_rows=%u",
              (int)r, (unsigned long long)rows, (unsigned long long)cols, box_rows);
  return SM3_OK;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per kernel instantiation and device: the launch-bound shapes pay
// ~3 us of host time for every redundant call (5 kernels per step)
#define SM3_SMEM_ATTR_ONCE(kernel, bytes)                                                                    \
  do {                                                                                                       \
    static int done_dev[16] = {0};                                                                           \
    int dev_ = 0;                                                                                            \
    cudaGetDevice(&dev_);                                                                                    \
    if (dev_ < 0 || dev_ >= 16 || !done_dev[dev_]) {                                                         \
      SM3_CHECK_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))); \
      if (dev_ >= 0 && dev_ < 16) done_dev[dev_] = 1;                                                        \
    }                                                                                                        \
  } while (0)

// ------------------------------------------------------------------------------------------------
// bring-up probe: one 128 x n x k MMA chain, every operand source/layout combination the real kernels use
// ------------------------------------------------------------------------------------------------
template <bool A_TMEM, bool B_MN>
__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const __nv_bfloat16* __restrict__ a_global, float* __restrict__ c, int n, int k) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t sA = base;                        // k/64 panels of [128 x 128 B]
  const uint32_t sB = base + 4 * 16384;            // K-major: k/64 panels [n x 128 B]; MN-major: n/64 panels [k x 128 B]
  const uint32_t bars = sB + 4 * 32768;
  const uint32_t bar_full = bars, bar_done = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + (bars - base) + 16);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    mbar_init(bar_full, 1);
    mbar_init(bar_done, 1);
    fence_barrier_init();
  }
  if (warp == 0) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();        // PDL: everything above overlapped the preceding kernel's tail; global memory from here on
  pdl_launch();
  const uint32_t tmem_d = tmem, tmem_a = tmem + 256;

  if (threadIdx.x == 0) {
    const uint32_t a_bytes = A_TMEM ? 0u : (uint32_t)(128 * k * 2);
    mbar_expect_tx(bar_full, a_bytes + (uint32_t)(n * k * 2));
    if (!A_TMEM)
      for (int pnl = 0; pnl < k / 64; ++pnl) tma_load_2d(sA + pnl * 16384, &tmap_a, bar_full, pnl * 64, 0);
    if (!B_MN)
      for (int pnl = 0; pnl < k / 64; ++pnl) tma_load_2d(sB + pnl * (n * 128), &tmap_b, bar_full, pnl * 64, 0);
    else
      for (int pnl = 0; pnl < n / 64; ++pnl) tma_load_2d(sB + pnl * (k * 128), &tmap_b, bar_full, pnl * 64, 0);
  }
  if (A_TMEM) {
    // thread = row; 32 words (64 bf16) per tcgen05.st; element k at column k/2, even k in the low half
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a_global + (size_t)threadIdx.x * k);
    for (int ch = 0; ch < k / 64; ++ch) {
      uint32_t r[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) r[i] = src[ch * 32 + i];
      tmem_st_x32(tmem_a + ((uint32_t)(warp * 32) << 16) + ch * 32, r);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (threadIdx.x == 0) {
    mbar_wait(bar_full, 0);
    tc_fence_after();
    const uint32_t idesc = make_idesc_bf16(128, n, 0, B_MN ? 1 : 0);
    for (int ks = 0; ks < k / 16; ++ks) {
      uint64_t bdesc;
      if (!B_MN) bdesc = make_smem_desc(sB + (ks >> 2) * (n * 128) + (ks & 3) * 32, 16, 1024);
      else bdesc = make_smem_desc(sB + ks * 2048, (uint32_t)(k * 128), 1024);
      if (A_TMEM) {
        umma_ts(tmem_d, tmem_a + ks * 8, bdesc, idesc, ks > 0);
      } else {
        const uint64_t adesc = make_smem_desc(sA + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024);
        umma_ss(tmem_d, adesc, bdesc, idesc, ks > 0);
      }
    }
    umma_commit(bar_done);
  }
  __syncwarp();
  mbar_wait(bar_done, 0);
  tc_fence_after();
  for (int ch = 0; ch < n / 32; ++ch) {
    uint32_t r[32];
    tmem_ld_x32(tmem_d + ((uint32_t)(warp * 32) << 16) + ch * 32, r);
    tmem_ld_wait(r);
#pragma unroll
    for (int i = 0; i < 32; ++i) c[(size_t)threadIdx.x * n + ch * 32 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
  (void)lane;
}

// ------------------------------------------------------------------------------------------------
// bring-up microbenchmark: dispatch rate of tcgen05.mma (M = 128, K = 16) for a given N and operand source, on an
// otherwise idle SM.  One thread issues `count` MMAs back to back into one accumulator; out[0] = cycles spent issuing,
// out[1] = cycles until the commit fires (all done).  Operand contents are irrelevant (whatever smem / TMEM hold).
// `ldtm_warps` > 0 adds that many warps streaming tcgen05.ld over the accumulator meanwhile (TMEM read contention).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(288, 1)
umma_rate_kernel(int n, int a_tmem_mode, int count, int ldtm_warps, long long* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t sA = base, sB = base + 16384, bars = sB + 65536;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + (bars - base) + 16);
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(bars, 1); fence_barrier_init(); stop = 0; }
  if (warp == 0) { tmem_alloc(smem_u32(tmem_slot), 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();        // PDL: everything above overlapped the preceding kernel's tail; global memory from here on
  pdl_launch();
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_bf16(128, n, 0, 0);
    const uint64_t bdesc = make_smem_desc(sB, 16, 1024);
    const uint64_t adesc = make_smem_desc(sA, 16, 1024);
    const long long t0 = clock64();
    for (int i = 0; i < count; ++i) {
      if (a_tmem_mode) umma_ts(tmem, tmem + 256 + (i & 3) * 8, bdesc, idesc, 1u);
      else umma_ss(tmem, adesc, bdesc, idesc, 1u);
    }
    const long long t1 = clock64();
    umma_commit(bars);
    mbar_wait(bars, 0);
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
    stop = 1;
  } else if (warp >= 1 && warp <= ldtm_warps) {
    long long reads = 0;
    uint32_t r[32];
    while (!stop) {
      tmem_ld_x32(tmem + ((uint32_t)((warp & 3) * 32) << 16) + 384, r);
      tmem_ld_wait(r);
      ++reads;
      if (r[0] == 0x12345678u && reads < 0) break;    // keep the loads alive
    }
    if ((threadIdx.x & 31) == 0) out[2 + warp] = reads;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int DP> struct FwdCfg {
  static constexpr int BN = 128;
  static constexpr uint32_t PANEL = BN * 128;          // 16 KB: 128 rows x 64 bf16
  static constexpr uint32_t STAGE = DP * PANEL;
  static constexpr int NSTAGE = DP == 4 ? 3 : 4;       // 192 KB / 192 KB / 128 KB / 64 KB
  static constexpr int NS = 3;                         // TMEM S stages at columns 128, 256, 384 (A at [0, 32*DP))
  static constexpr uint32_t SMEM = NSTAGE * STAGE + 1024 /*align*/ + 256 /*barriers*/ + 3 * 512 /*xsum*/;
};

// POLY: how many of every 8 exponentials run on the FMA pipes (ex2_fma) instead of MUFU.
template <int DP, int NG, int POLY>
__global__ void __launch_bounds__(64 + 256 * NG, 1)
infonce_tc_fwd_kernel(const __grid_constant__ CUtensorMap tmap_cols, TcParams p) {
  using C = FwdCfg<DP>;
  constexpr int BN = C::BN, NSTAGE = C::NSTAGE, NS = C::NS;
  constexpr int D = 64 * DP;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t sB = base;
  const uint32_t bars = sB + NSTAGE * C::STAGE;
  auto bar_full = [&](int i) { return bars + 8u * i; };
  auto bar_empty = [&](int i) { return bars + 8u * (NSTAGE + i); };
  auto bar_sfull = [&](int i) { return bars + 8u * (2 * NSTAGE + i); };
  auto bar_sempty = [&](int i) { return bars + 8u * (2 * NSTAGE + NS + i); };
  const uint32_t bar_aready = bars + 8u * (2 * NSTAGE + 2 * NS);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + (bars - base) + 8u * (2 * NSTAGE + 2 * NS + 1));
  float* xsum = reinterpret_cast<float*>(base_ptr + (bars - base) + 256);   // [2*NG - 1][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned long long bar_limit = 2000000000ull + p.wait_timeout_ns;   // + cross-rank bound in fused-exchange mode
  const int r0 = blockIdx.x * kBM;
  const int split = blockIdx.y;
  const int t_begin = split * p.tiles_per_split;
  const int t_end = min(p.col_tiles, t_begin + p.tiles_per_split);
  const int n_tiles = t_end - t_begin;
  // Column tiles are visited in a per-CTA rotated order.  All row blocks of a wave would otherwise sweep the SAME
  // 32-64 KB tile at the same time and serialise on the few L2 slices holding it (measured: the forward kernel was
  // bimodal, 1.7 ms or 3.5-6 ms at cfg4, depending on whether the CTAs happened to run in lockstep).
  const int rot = tile_rotation(p, t_begin, n_tiles);
  auto tile_of = [&](int it) {
    int t = it + rot;
    t = t_begin + (t >= n_tiles ? t - n_tiles : t);
    if (t >= p.skip_a) t += p.skip_len;        // skip mode: hop over the column tiles owned by this rank
    if (t >= p.skip_b) t += p.skip_len;
    return t;
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(bar_full(i), 1); mbar_init(bar_empty(i), 1); }
    for (int i = 0; i < NS; ++i) { mbar_init(bar_sfull(i), 1); mbar_init(bar_sempty(i), 8); }
    mbar_init(bar_aready, 8 * NG);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();        // PDL: everything above overlapped the preceding kernel's tail; global memory from here on
  pdl_launch();

  if (warp == 0) {
    // =========================== TMA producer (warp-uniform loop, one elected lane issues) ===========
    if (elect_one()) prefetch_tensormap(&tmap_cols);
    bool remote_ready = (p.wait_flags == nullptr);
    for (int it = 0; it < n_tiles; ++it) {
      const int s = it % NSTAGE;
      const uint32_t ph = (uint32_t)(it / NSTAGE) & 1u;
      const int tile = tile_of(it);
      if (!remote_ready && !tile_is_local(p, tile)) {     // fused exchange: the peers' rows must have landed
        if (elect_one()) {
          peer_flags_wait_all(p.wait_flags, p.wait_world, p.wait_channel, p.wait_epoch, p.wait_timeout_ns);
          fence_proxy_async_global();
        }
        __syncwarp();
        remote_ready = true;
      }
      mbar_wait(bar_empty(s), ph ^ 1u, bar_limit);
      if (elect_one()) {
        mbar_expect_tx(bar_full(s), C::STAGE);
        const int row = tile * BN;
#pragma unroll
        for (int pnl = 0; pnl < DP; ++pnl)
          tma_load_2d(sB + s * C::STAGE + pnl * C::PANEL, &tmap_cols, bar_full(s), pnl * 64, row);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (warp-uniform loop, one elected lane issues) ============
    mbar_wait(bar_aready, 0, bar_limit);
    tc_fence_after();
    constexpr uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
    constexpr uint32_t dhi = smem_desc_hi(1024);
    for (int it = 0; it < n_tiles; ++it) {
      const int s = it % NSTAGE, as = it % NS;
      const uint32_t ph = (uint32_t)(it / NSTAGE) & 1u, aph = (uint32_t)(it / NS) & 1u;
      mbar_wait(bar_sempty(as), aph ^ 1u, bar_limit);
      mbar_wait(bar_full(s), ph, bar_limit);
      tc_fence_after();
      SM3_TR(0, it);
      const uint32_t d_tmem = tmem + 128u + (uint32_t)as * 128u;
      const uint32_t lo0 = smem_desc_lo(sB + s * C::STAGE, 16);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4 * DP; ++ks)
          umma_ts(d_tmem, tmem + ks * 8, desc64(lo0 + (((ks >> 2) * C::PANEL + (ks & 3) * 32) >> 4), dhi), idesc,
                  ks > 0);
        umma_commit(bar_empty(s));
        umma_commit(bar_sfull(as));
      }
      __syncwarp();
      SM3_TR(2, it);
    }
  } else {
    // =========================== softmax warps ===========================
    // NG groups of 8 warps; group g owns the tiles with it % NG == g, so consecutive tiles are processed by
    // different warps of the same SM sub-partition and their TMEM-load / barrier latencies overlap.
    const int q = warp & 3;                    // TMEM lane quadrant this warp may touch
    const int grp = (warp - 2) >> 3;           // softmax group
    const int half = ((warp - 2) >> 2) & 1;    // which 64 of the tile's 128 columns
    const int combo = grp * 2 + half;
    const int row_in_tile = q * 32 + lane;
    const int l = r0 + row_in_tile;
    const bool valid = l < p.m_rows;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;

    // ---- stage this CTA's 128 rows into TMEM as the A operand (packed bf16 pairs) ----
#pragma unroll
    for (int ch = 0; ch < DP; ++ch) {
      if ((ch % (2 * NG)) == combo) {
        uint32_t r[32];
        if (valid) {
          const uint4* src = reinterpret_cast<const uint4*>(p.z_rows + (size_t)l * D + ch * 64);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint4 v = __ldg(src + i);
            r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = 0u;
        }
        tmem_st_x32(tmem + lane_addr + ch * 32, r);
      }
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_aready);

    const int g = valid ? global_row(l, p.n_local, p.pair_offset, p.n_global) : -1;
    const int pj = valid ? positive_of(g, p.n_global) : -1;
    const float c2 = p.c2;
    float sum0 = 0.f, sum1 = 0.f, sum2 = 0.f, sum3 = 0.f;
    float posval = 0.f;
    bool found = false;

    // (Register double-buffering of the next tile's TMEM loads was tried and is slower: it delays the release of
    // the S stage by a whole tile of exponentials and starves the MMA warp.)
    for (int it = grp; it < n_tiles; it += NG) {
      const int as = it % NS;
      mbar_wait(bar_sfull(as), (uint32_t)(it / NS) & 1u, bar_limit);
      tc_fence_after();
      if (warp == 2 && lane == 0) SM3_TR(3, it);
      const uint32_t taddr = tmem + lane_addr + 128u + (uint32_t)as * 128u + (uint32_t)half * 64u;
      uint32_t a[32], b[32];
      tmem_ld_x32(taddr, a);
      tmem_ld_x32(taddr + 32, b);
      tmem_ld_wait(a);
      tmem_ld_wait(b);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_sempty(as));      // S stage is in registers: hand it back to the MMA warp
      if (warp == 2 && lane == 0) SM3_TR(4, it);

      const int cb = tile_of(it) * BN + half * 64;
      const bool need = (cb + 64 > p.m_cols) ||
                        (valid && ((unsigned)(g - cb) < 64u || (unsigned)(pj - cb) < 64u));
      if (!__any_sync(0xffffffffu, need)) {
#pragma unroll
        for (int i = 0; i < 64; i += 4) {
          const uint32_t* v = i < 32 ? a : b;
          const float x0 = fmaf(__uint_as_float(v[(i) & 31]), c2, -c2);
          const float x1 = fmaf(__uint_as_float(v[(i + 1) & 31]), c2, -c2);
          const float x2 = fmaf(__uint_as_float(v[(i + 2) & 31]), c2, -c2);
          const float x3 = fmaf(__uint_as_float(v[(i + 3) & 31]), c2, -c2);
          sum0 += ((i & 7) < POLY) ? ex2_fma(x0) : ex2(x0);
          sum1 += (((i + 1) & 7) < POLY) ? ex2_fma(x1) : ex2(x1);
          sum2 += (((i + 2) & 7) < POLY) ? ex2_fma(x2) : ex2(x2);
          sum3 += (((i + 3) & 7) < POLY) ? ex2_fma(x3) : ex2(x3);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          const float s = __uint_as_float(i < 32 ? a[i & 31] : b[i & 31]);
          const int col = cb + i;
          const bool is_pos = (col == pj);
          if (is_pos) { posval = s * p.inv_T; found = true; }
          if (col < p.m_cols && col != g && !is_pos) sum0 += ex2(fmaf(s, c2, -c2));
        }
      }
      if (warp == 2 && lane == 0) SM3_TR(5, it);
    }
    float total = (sum0 + sum1) + (sum2 + sum3);
    if (found) p.pos[l] = posval;
    if (combo != 0) xsum[(combo - 1) * 128 + row_in_tile] = total;
    named_bar_sync(1, 256 * NG);
    if (combo == 0 && valid) {
#pragma unroll
      for (int c = 0; c < 2 * NG - 1; ++c) total += xsum[c * 128 + row_in_tile];
      p.partial[(size_t)split * p.m_rows + l] = total;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// forward, 256 rows per CTA: every B tile fetched from L2 feeds TWO S-MMAs (row blocks r0 and r0+128).
// The 128-row kernel above moves 64 KB of L2 data per 128x128x256 tile; over 148 SMs that is ~6.7 KB/clk, right at
// the chip's L2 throughput cap (~6.3 KB/clk), which is why it was slow and erratic at cfg4.  Here the traffic halves.
// TMEM: A0 [0,128) | A1 [128,256) | S(row block 0) [256,384) | S(row block 1) [384,512); the two S stages ping-pong
// between the MMA warp and the 8 softmax warps exactly like FA-style two-Q-tile kernels.
// ------------------------------------------------------------------------------------------------
// kPush: two extra warps (10, 11) push this rank's rows to the peers while the first column tiles -- this rank's own --
// are being computed (multi-rank fused exchange, owner-ordered tiles; see TcParams).
template <int DP, int POLY, bool kPush>
__global__ void __launch_bounds__(kPush ? 384 : 320, 1)
infonce_tc_fwd2_kernel(const __grid_constant__ CUtensorMap tmap_cols, TcParams p) {
  using C = FwdCfg<DP>;
  constexpr int BN = C::BN, NSTAGE = C::NSTAGE;
  constexpr int D = 64 * DP;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t sB = base;
  const uint32_t bars = sB + NSTAGE * C::STAGE;
  auto bar_full = [&](int i) { return bars + 8u * i; };
  auto bar_empty = [&](int i) { return bars + 8u * (NSTAGE + i); };
  auto bar_sfull = [&](int i) { return bars + 8u * (2 * NSTAGE + i); };
  auto bar_sempty = [&](int i) { return bars + 8u * (2 * NSTAGE + 2 + i); };
  const uint32_t bar_aready = bars + 8u * (2 * NSTAGE + 4);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + (bars - base) + 8u * (2 * NSTAGE + 5));
  float* xsum = reinterpret_cast<float*>(base_ptr + (bars - base) + 256);   // [2 row blocks][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) SM3_TR(7, 0);                                      // trace: CTA start
  const unsigned long long bar_limit = 2000000000ull + p.wait_timeout_ns;   // + cross-rank bound in fused-exchange mode
  const int r0 = blockIdx.x * 256;
  const int split = blockIdx.y;
  const int t_begin = split * p.tiles_per_split;
  const int t_end = min(p.col_tiles, t_begin + p.tiles_per_split);
  const int n_tiles = t_end - t_begin;
  const int own_per = p.own_order ? p.own_L / p.own_S : 1;
  const int rot = p.own_order ? (int)((blockIdx.x * 37u + blockIdx.y * 11u) % (unsigned)own_per)
                              : tile_rotation(p, t_begin, n_tiles);
  auto tile_of = [&](int it) {
    if (p.own_order) { int o; return owner_tile(p, it, split, rot, &o); }
    int t = it + rot;
    t = t_begin + (t >= n_tiles ? t - n_tiles : t);
    if (t >= p.skip_a) t += p.skip_len;        // skip mode: hop over the column tiles owned by this rank
    if (t >= p.skip_b) t += p.skip_len;
    return t;
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(bar_full(i), 1); mbar_init(bar_empty(i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_sfull(i), 1); mbar_init(bar_sempty(i), 8); }
    mbar_init(bar_aready, 8);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();        // PDL: everything above overlapped the preceding kernel's tail; global memory from here on
  pdl_launch();
  constexpr uint32_t kColA1 = 128, kColS = 256;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (elect_one()) prefetch_tensormap(&tmap_cols);
    bool remote_ready = (p.wait_flags == nullptr);
    int ready_pos = 0;                                    // owner-ordered mode: owners [0, ready_pos] have landed
    for (int it = 0; it < n_tiles; ++it) {
      const int s = it % NSTAGE;
      const uint32_t ph = (uint32_t)(it / NSTAGE) & 1u;
      const int tile = tile_of(it);
      if (p.own_order) {
        const int pos = it / (2 * own_per);
        if (pos > ready_pos) {                            // first tile of the next owner: its rows must have landed
          int o;
          owner_tile(p, it, split, rot, &o);
          if (elect_one()) {
            peer_flag_wait_one(p.wait_flags, p.wait_channel, o, p.wait_epoch, p.wait_timeout_ns);
            fence_proxy_async_global();
          }
          __syncwarp();
          ready_pos = pos;
        }
      } else if (!remote_ready && !tile_is_local(p, tile)) {     // fused exchange: the peers' rows must have landed
        if (elect_one()) {
          peer_flags_wait_all(p.wait_flags, p.wait_world, p.wait_channel, p.wait_epoch, p.wait_timeout_ns);
          fence_proxy_async_global();
        }
        __syncwarp();
        remote_ready = true;
      }
      mbar_wait(bar_empty(s), ph ^ 1u, bar_limit);
      if (elect_one()) {
        mbar_expect_tx(bar_full(s), C::STAGE);
        const int row = tile * BN;
#pragma unroll
        for (int pnl = 0; pnl < DP; ++pnl)
          tma_load_2d(sB + s * C::STAGE + pnl * C::PANEL, &tmap_cols, bar_full(s), pnl * 64, row);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    mbar_wait(bar_aready, 0, bar_limit);
    tc_fence_after();
    SM3_TR(7, 1);                                                          // trace: rows staged, first MMA
    constexpr uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
    constexpr uint32_t dhi = smem_desc_hi(1024);
    for (int it = 0; it < n_tiles; ++it) {
      const int s = it % NSTAGE;
      const uint32_t sph = (uint32_t)it & 1u;          // each S stage is used once per tile
      mbar_wait(bar_full(s), (uint32_t)(it / NSTAGE) & 1u, bar_limit);
      SM3_TR(0, it);
      const uint32_t lo0 = smem_desc_lo(sB + s * C::STAGE, 16);
#pragma unroll
      for (int rb = 0; rb < 2; ++rb) {
        mbar_wait(bar_sempty(rb), sph ^ 1u, bar_limit);
        tc_fence_after();
        SM3_TR(8 + rb, it);                                                // trace: S stage rb free, issue starts
        const uint32_t d_tmem = tmem + kColS + (uint32_t)rb * 128u;
        const uint32_t a_tmem = tmem + (uint32_t)rb * kColA1;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < 4 * DP; ++ks)
            umma_ts(d_tmem, a_tmem + ks * 8, desc64(lo0 + (((ks >> 2) * C::PANEL + (ks & 3) * 32) >> 4), dhi), idesc,
                    ks > 0);
          if (rb == 1) umma_commit(bar_empty(s));
          umma_commit(bar_sfull(rb));
        }
        __syncwarp();
        SM3_TR(1 + rb, it);
      }
    }
    SM3_TR(7, 2);
  } else if (kPush && warp >= 10) {
    // =========================== row pushers (first-wave CTAs only) ===========================
    const unsigned cta = blockIdx.y * gridDim.x + blockIdx.x;
    if (p.push_src != nullptr && cta < (unsigned)p.push_ctas) {
      const int t = (int)threadIdx.x - 320;
      const long total = (long)p.m_rows * p.push_vpr;
      const long per_cta = (total + p.push_ctas - 1) / p.push_ctas;
      const long begin = (long)cta * per_cta, end = min(total, begin + per_cta);
      const int me = p.pair_offset / p.n_local;
      for (int j = 1; j < p.wait_world; ++j) {
        const int dst = (me + j) % p.wait_world;          // rank+1 first: it visits OUR columns first after its own
        uint4* out = reinterpret_cast<uint4*>(p.push_dst.p[dst]);
        for (long i = begin + t; i < end; i += 64) {
          const int lrow = (int)(i / p.push_vpr);
          const int v = (int)(i - (long)lrow * p.push_vpr);
          const int grow = global_row(lrow, p.n_local, p.pair_offset, p.n_global);
          out[(long)grow * p.push_vpr + v] = __ldg(p.push_src + i);
        }
        named_bar_sync(2, 64);                            // all 64 pushers' stores precede thread 0's fence + ticket
        if (t == 0) {
          __threadfence_system();
          const unsigned k = atomicAdd(p.push_counter + dst, 1u);
          if (k == (unsigned)p.push_ctas - 1u) {          // every slice of this destination is visible: release its flag
            p.push_counter[dst] = 0u;
            __threadfence_system();
            unsigned* f = reinterpret_cast<unsigned*>(p.push_flags.p[dst]) + p.wait_channel * 16 + me;
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(p.wait_epoch) : "memory");
          }
        }
      }
    }
  } else {
    // =========================== softmax warps ===========================
    const int q = warp & 3;
    const int half = ((warp - 2) >> 2) & 1;
    const int row_in_tile = q * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    int l[2], g[2], pj[2];
    bool valid[2];
#pragma unroll
    for (int rb = 0; rb < 2; ++rb) {
      l[rb] = r0 + rb * 128 + row_in_tile;
      valid[rb] = l[rb] < p.m_rows;
      g[rb] = valid[rb] ? global_row(l[rb], p.n_local, p.pair_offset, p.n_global) : -1;
      pj[rb] = valid[rb] ? positive_of(g[rb], p.n_global) : -1;
    }
    // ---- stage both row blocks into TMEM as A operands ----
#pragma unroll
    for (int rb = 0; rb < 2; ++rb) {
#pragma unroll
      for (int ch = 0; ch < DP; ++ch) {
        if ((ch & 1) == half) {
          uint32_t r[32];
          if (valid[rb]) {
            const uint4* src = reinterpret_cast<const uint4*>(p.z_rows + (size_t)l[rb] * D + ch * 64);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint4 v = __ldg(src + i);
              r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i) r[i] = 0u;
          }
          tmem_st_x32(tmem + lane_addr + rb * kColA1 + ch * 32, r);
        }
      }
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_aready);

    const float c2 = p.c2;
    float sum[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
    float posval[2] = {0.f, 0.f};
    bool found[2] = {false, false};

    for (int it = 0; it < n_tiles; ++it) {
      const uint32_t sph = (uint32_t)it & 1u;
      const int cb = tile_of(it) * BN + half * 64;
#pragma unroll
      for (int rb = 0; rb < 2; ++rb) {
        mbar_wait(bar_sfull(rb), sph, bar_limit);
        tc_fence_after();
        if (warp == 2 && lane == 0) SM3_TR(3 + 2 * rb, it);
        const uint32_t taddr = tmem + lane_addr + kColS + (uint32_t)rb * 128u + (uint32_t)half * 64u;
        uint32_t a[32], b[32];
        tmem_ld_x32(taddr, a);
        tmem_ld_x32(taddr + 32, b);
        tmem_ld_wait(a);
        tmem_ld_wait(b);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_sempty(rb));
        if (warp == 2 && lane == 0) SM3_TR(4 + 2 * rb, it);
        const bool need = (cb + 64 > p.m_cols) ||
                          (valid[rb] && ((unsigned)(g[rb] - cb) < 64u || (unsigned)(pj[rb] - cb) < 64u));
        if (!__any_sync(0xffffffffu, need)) {
#pragma unroll
          for (int i = 0; i < 64; i += 4) {
            const uint32_t* v = i < 32 ? a : b;
            const float x0 = fmaf(__uint_as_float(v[(i) & 31]), c2, -c2);
            const float x1 = fmaf(__uint_as_float(v[(i + 1) & 31]), c2, -c2);
            const float x2 = fmaf(__uint_as_float(v[(i + 2) & 31]), c2, -c2);
            const float x3 = fmaf(__uint_as_float(v[(i + 3) & 31]), c2, -c2);
            sum[rb][0] += ((i & 7) < POLY) ? ex2_fma(x0) : ex2(x0);
            sum[rb][1] += (((i + 1) & 7) < POLY) ? ex2_fma(x1) : ex2(x1);
            sum[rb][2] += (((i + 2) & 7) < POLY) ? ex2_fma(x2) : ex2(x2);
            sum[rb][3] += (((i + 3) & 7) < POLY) ? ex2_fma(x3) : ex2(x3);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 64; ++i) {
            const float sv = __uint_as_float(i < 32 ? a[i & 31] : b[i & 31]);
            const int col = cb + i;
            const bool is_pos = (col == pj[rb]);
            if (is_pos) { posval[rb] = sv * p.inv_T; found[rb] = true; }
            if (col < p.m_cols && col != g[rb] && !is_pos) sum[rb][0] += ex2(fmaf(sv, c2, -c2));
          }
        }
      }
    }
#pragma unroll
    for (int rb = 0; rb < 2; ++rb) {
      if (found[rb]) p.pos[l[rb]] = posval[rb];
      const float total = (sum[rb][0] + sum[rb][1]) + (sum[rb][2] + sum[rb][3]);
      if (half == 1) xsum[rb * 128 + row_in_tile] = total;
      named_bar_sync(1, 256);
      if (half == 0 && valid[rb]) p.partial[(size_t)split * p.m_rows + l[rb]] = total + xsum[rb * 128 + row_in_tile];
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) SM3_TR(7, 4);                                      // trace: CTA end
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// symmetric forward (single rank, rows == columns): S = Z Z^T is symmetric, so only the column tiles J >= 2R of every
// 256-row pair R are computed; a tile with J > 2R + 1 gives its row sums to the rows of R and its COLUMN sums to the rows
// of tile J (what the transposed tile would have produced): half the MMAs and half the exponentials of the kernel above.
// Work list, workspace layout and the fold: see kSymFlag in common.cuh.  Same warp roles / TMEM / smem ring as the
// 256-row kernel; differences:
//   * a CTA walks a contiguous piece of the flat (R, J) list and re-stages its A rows when R changes;
//   * the softmax warps read S in the 16x256b fragment layout (thread = 4 rows x 16 columns of its 32 x 64 slice), add
//     their rows per column while they accumulate the row sums (1 extra FADD per logit), finish the column sums with a
//     3-step butterfly over the 8 lanes that share a column (14 shuffles per tile), fold the four lane quadrants through
//     shared memory in a fixed order and write 128 floats per tile;
//   * the row sums need their 2-step cross-lane reduction only once per (CTA, R) segment.
// ------------------------------------------------------------------------------------------------
template <int DP> struct SymCfg {
  static constexpr uint32_t SMEM = FwdCfg<DP>::NSTAGE * FwdCfg<DP>::STAGE + 1024 /*align*/ + 256 /*barriers*/ +
                                   3072 /*xsum [3][2][128]*/ + 6144 /*cbuf [3][512]*/;
};
__host__ __device__ __forceinline__ void sym_decode(long f, int T, int P, int& R, int& off) {
  const int i = (int)(f / (T + 2));
  const int rem = (int)(f - (long)i * (T + 2));
  const int len_i = T - 2 * i;
  if (rem < len_i) { R = i; off = rem; } else { R = P - 1 - i; off = rem - len_i; }
}

// NQ = softmax warps per lane quadrant (2 or 4): each owns 128 / NQ columns of every S tile.  4 (16 softmax warps, 576
// threads, <= 112 registers) doubles the warps that hide each other's MUFU / TMEM / shuffle latencies.
template <int DP, int POLY, int NQ>
__global__ void __launch_bounds__(64 + 128 * NQ, 1)
infonce_tc_fwdsym_kernel(const __grid_constant__ CUtensorMap tmap_cols, TcParams p) {
  using C = FwdCfg<DP>;
  constexpr int BN = C::BN, NSTAGE = C::NSTAGE;
  constexpr int D = 64 * DP;
  constexpr int NW = 4 * NQ;             // softmax warps
  constexpr int CW = 128 / NQ;           // S-tile columns per softmax warp
  constexpr int KS = CW / 8;             // 8-column groups per warp: registers 4 k .. 4 k + 3 of a 16x256b load
  const int T = p.sym_T, P = T >> 1;
  const long f0 = (long)blockIdx.x * p.sym_tpc;
  const long f1 = min(p.sym_W, f0 + p.sym_tpc);
  if (f0 >= f1) return;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t sB = base;
  const uint32_t bars = sB + NSTAGE * C::STAGE;
  auto bar_full = [&](int i) { return bars + 8u * i; };
  auto bar_empty = [&](int i) { return bars + 8u * (NSTAGE + i); };
  auto bar_sfull = [&](int i) { return bars + 8u * (2 * NSTAGE + i); };
  auto bar_sempty = [&](int i) { return bars + 8u * (2 * NSTAGE + 2 + i); };
  const uint32_t bar_aready = bars + 8u * (2 * NSTAGE + 4);
  auto bar_col = [&](int i) { return bars + 8u * (2 * NSTAGE + 5 + i); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + (bars - base) + 8u * (2 * NSTAGE + 8));
  float* xsum = reinterpret_cast<float*>(base_ptr + (bars - base) + 256);          // [NQ - 1][2 row blocks][128]
  float* cbuf = xsum + 768;                                                         // [3 ring slots][NW warps][CW]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned long long bar_limit = 2000000000ull;
  if (threadIdx.x == 0) SM3_TR(7, 0);

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(bar_full(i), 1); mbar_init(bar_empty(i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_sfull(i), 1); mbar_init(bar_sempty(i), NW); }
    mbar_init(bar_aready, NW);
    for (int i = 0; i < 3; ++i) mbar_init(bar_col(i), NW);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();
  pdl_launch();
  constexpr uint32_t kColA1 = 128, kColS = 256;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (elect_one()) prefetch_tensormap(&tmap_cols);
    int it = 0;
    for (long f = f0; f < f1;) {
      int R, off;
      sym_decode(f, T, P, R, off);
      const int cnt = (int)min((long)(T - 2 * R - off), f1 - f);
      for (int t = 0; t < cnt; ++t, ++it) {
        const int s = it % NSTAGE;
        const uint32_t ph = (uint32_t)(it / NSTAGE) & 1u;
        mbar_wait(bar_empty(s), ph ^ 1u, bar_limit);
        if (elect_one()) {
          mbar_expect_tx(bar_full(s), C::STAGE);
          const int row = (2 * R + off + t) * BN;
#pragma unroll
          for (int pnl = 0; pnl < DP; ++pnl)
            tma_load_2d(sB + s * C::STAGE + pnl * C::PANEL, &tmap_cols, bar_full(s), pnl * 64, row);
        }
        __syncwarp();
      }
      f += cnt;
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    constexpr uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
    constexpr uint32_t dhi = smem_desc_hi(1024);
    int it = 0, seg = 0;
    for (long f = f0; f < f1; ++seg) {
      int R, off;
      sym_decode(f, T, P, R, off);
      const int cnt = (int)min((long)(T - 2 * R - off), f1 - f);
      mbar_wait(bar_aready, (uint32_t)seg & 1u, bar_limit);      // this row pair's A operands are in TMEM
      tc_fence_after();
      for (int t = 0; t < cnt; ++t, ++it) {
        const int s = it % NSTAGE;
        const uint32_t sph = (uint32_t)it & 1u;
        mbar_wait(bar_full(s), (uint32_t)(it / NSTAGE) & 1u, bar_limit);
        if (it == 0) SM3_TR(7, 1);
        SM3_TR(0, it);
        const uint32_t lo0 = smem_desc_lo(sB + s * C::STAGE, 16);
#pragma unroll
        for (int rb = 0; rb < 2; ++rb) {
          mbar_wait(bar_sempty(rb), sph ^ 1u, bar_limit);
          tc_fence_after();
          SM3_TR(8 + rb, it);
          const uint32_t d_tmem = tmem + kColS + (uint32_t)rb * 128u;
          const uint32_t a_tmem = tmem + (uint32_t)rb * kColA1;
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4 * DP; ++ks)
              umma_ts(d_tmem, a_tmem + ks * 8, desc64(lo0 + (((ks >> 2) * C::PANEL + (ks & 3) * 32) >> 4), dhi), idesc,
                      ks > 0);
            if (rb == 1) umma_commit(bar_empty(s));
            umma_commit(bar_sfull(rb));
          }
          __syncwarp();
          SM3_TR(1 + rb, it);
        }
      }
      f += cnt;
    }
    SM3_TR(7, 2);
  } else {
    // =========================== softmax warps ===========================
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;                              // column group of this warp: columns [cg * CW, +CW)
    const int t0 = lane & 3, t1 = lane >> 2;
    const int row_in_tile = q * 32 + lane;                       // staging view: thread <-> TMEM lane
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint32_t lane_hi = (uint32_t)(q * 32 + 16) << 16;
    const float c2 = p.c2;
    const uint64_t c2p = f2_pack(c2, c2), nc2p = f2_pack(-c2, -c2);
    const bool clamp = 2.0f * c2 > 125.0f;
    const int64_t M = p.m_rows;
    float* colpart = p.partial + (int64_t)p.sym_maxseg * M;
    int it = 0;
    // column sums are folded across the four lane quadrants one tile LATE (no CTA-wide rendezvous per tile): ring of 3
    // buffers, one mbarrier each (NW warp arrivals); pend_* describe the tile whose fold is outstanding
    int ncol = 0, pend_R = -1, pend_J = 0;
    auto fold_pending = [&]() {
      if (pend_R < 0) return;
      const int ci = (ncol - 1) % 3;
      mbar_wait(bar_col(ci), (uint32_t)((ncol - 1) / 3) & 1u, bar_limit);
      if (lane < 128 / NW) {                                     // NW warps x 128 / NW columns, quadrants in a fixed order
        const int c = (warp - 2) * (128 / NW) + lane;
        const float* src = cbuf + ci * 512 + (c / CW) * (4 * CW) + (c % CW);
        colpart[(int64_t)pend_R * M + (int64_t)pend_J * BN + c] = (src[0] + src[CW]) + (src[2 * CW] + src[3 * CW]);
      }
      pend_R = -1;
    };
    for (long f = f0; f < f1;) {
      int R, off;
      sym_decode(f, T, P, R, off);
      const int cnt = (int)min((long)(T - 2 * R - off), f1 - f);
      const int r0 = R * 256;
      // ---- stage this row pair into TMEM as A operands (all S-MMAs of the previous segment have completed: every
      //      softmax warp has waited for the last S tile) ----
#pragma unroll
      for (int rb = 0; rb < 2; ++rb) {
#pragma unroll
        for (int ch = 0; ch < DP; ++ch) {
          if ((ch % NQ) == cg) {
            uint32_t r[32];
            const uint4* src = reinterpret_cast<const uint4*>(p.z_rows + (size_t)(r0 + rb * 128 + row_in_tile) * D + ch * 64);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint4 v = __ldg(src + i);
              r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
            }
            tmem_st_x32(tmem + lane_addr + rb * kColA1 + ch * 32, r);
          }
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_aready);

      // packed (column b = 0 | b = 1) running row sums of the thread's 4 rows per row block
      uint64_t sum2[2][4];
#pragma unroll
      for (int rb = 0; rb < 2; ++rb)
#pragma unroll
        for (int sl = 0; sl < 4; ++sl) sum2[rb][sl] = 0ull;
      for (int t = 0; t < cnt; ++t, ++it) {
        const int J = 2 * R + off + t;
        const uint32_t sph = (uint32_t)it & 1u;
        const bool do_col = J > 2 * R + 1;
        uint64_t col2[KS];                                       // packed column sums: columns 8 k + 2 t0 + {0, 1}
        // TMEM loads run one half tile ahead of the exponentials: vb(rb) is in flight during the first half of row block
        // rb, and va of row block 1 (its S tile has been ready for a while: the MMA warp runs ahead) during the second
        // half of row block 0 -- only the first load of a tile is exposed.
        uint32_t va[4 * KS], vb[4 * KS];
        auto begin_rb = [&](int rb) {
          mbar_wait(bar_sfull(rb), sph, bar_limit);
          tc_fence_after();
          if (warp == 2 && lane == 0) SM3_TR(3 + 2 * rb, it);
          tmem_ld_16x256b(tmem + kColS + (uint32_t)rb * 128u + (uint32_t)(cg * CW) + lane_addr, va);   // rows t1, t1 + 8
        };
        begin_rb(0);
#pragma unroll
        for (int rb = 0; rb < 2; ++rb) {
          const uint32_t taddr = tmem + kColS + (uint32_t)rb * 128u + (uint32_t)(cg * CW);
          tmem_ld_wait(va);
          tmem_ld_16x256b(taddr + lane_hi, vb);                  // rows t1 + 16, t1 + 24: in flight during the first half
          int Jp = 2 * R + rb + P;                               // the tile that holds these rows' positives
          if (Jp >= T) Jp -= T;
          const bool special = (J == 2 * R + rb) || (J == Jp);
#pragma unroll
          for (int hv = 0; hv < 2; ++hv) {
            if (hv == 1) {
              tmem_ld_wait(vb);
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar_sempty(rb));        // S stage free (the MMA warp runs a tile ahead anyway)
              if (warp == 2 && lane == 0) SM3_TR(4 + 2 * rb, it);
              if (rb == 0) begin_rb(1);                          // va is dead: prefetch row block 1's first half
            }
            const uint32_t* v = hv == 0 ? va : vb;
            if (!special) {
              uint64_t e[2 * KS];
#pragma unroll
              for (int k = 0; k < KS; ++k) {
                e[2 * k] = f2_fma(f2_pack_u(v[4 * k], v[4 * k + 1]), c2p, nc2p);
                e[2 * k + 1] = f2_fma(f2_pack_u(v[4 * k + 2], v[4 * k + 3]), c2p, nc2p);
              }
              // POLY of every 8 exponentials on the FMA pipes: whole pairs (4: every other pair; 2: every fourth pair)
#pragma unroll
              for (int i = 0; i < 2 * KS; ++i)
                e[i] = ((POLY >= 4 && (i & 1) == 0) || (POLY >= 2 && POLY < 4 && (i & 3) == 0)) ? f2_ex2_fma(e[i], clamp) : f2_ex2(e[i]);
#pragma unroll
              for (int k = 0; k < KS; ++k) {
                sum2[rb][2 * hv] = f2_add(sum2[rb][2 * hv], e[2 * k]);
                sum2[rb][2 * hv + 1] = f2_add(sum2[rb][2 * hv + 1], e[2 * k + 1]);
                const uint64_t cs = f2_add(e[2 * k], e[2 * k + 1]);
                col2[k] = (rb == 0 && hv == 0) ? cs : f2_add(col2[k], cs);
              }
            } else {
              // the tile holds the diagonal or the positives of these rows: mask per element
              const int rbase = r0 + rb * 128 + q * 32 + t1 + 16 * hv;
              const int cbase = J * BN + cg * CW + 2 * t0;
#pragma unroll
              for (int k = 0; k < KS; ++k) {
                uint64_t cs = 0ull;
#pragma unroll
                for (int a8 = 0; a8 < 2; ++a8) {
                  const int row = rbase + 8 * a8;
                  const int pj = positive_of(row, p.n_global);
                  float ev[2];
#pragma unroll
                  for (int bb = 0; bb < 2; ++bb) {
                    const int col = cbase + 8 * k + bb;
                    const float sv = __uint_as_float(v[4 * k + 2 * a8 + bb]);
                    const bool is_pos = (col == pj);
                    if (is_pos && pj > row) {                    // S is symmetric: one read serves both rows of the pair
                      const float pv = sv * p.inv_T;
                      p.pos[row] = pv;
                      p.pos[pj] = pv;
                    }
                    ev[bb] = (is_pos || col == row) ? 0.f : ex2(fmaf(sv, c2, -c2));
                  }
                  const uint64_t e = f2_pack(ev[0], ev[1]);
                  sum2[rb][2 * hv + a8] = f2_add(sum2[rb][2 * hv + a8], e);
                  cs = f2_add(cs, e);
                }
                col2[k] = (rb == 0 && hv == 0) ? cs : f2_add(col2[k], cs);
              }
            }
          }
        }
        if (warp == 2 && lane == 0) SM3_TR(10, it);
        fold_pending();                                          // the previous column-sum tile: every warp arrived long ago
        if (do_col) {
          // column sums over this warp's 32 rows x 2 row blocks: butterfly over the lanes t1 = lane / 4 that share a
          // column (lane bit 4 <-> k bit 2 or 1, ...); each step halves the values a lane carries
          float ca[2 * KS];
#pragma unroll
          for (int k = 0; k < KS; ++k) f2_unpack(col2[k], ca[2 * k], ca[2 * k + 1]);
          const bool h4 = (lane & 16) != 0, h2 = (lane & 8) != 0, h1 = (lane & 4) != 0;
          float* cb = cbuf + (ncol % 3) * 512 + (cg * 4 + q) * CW;
          if constexpr (KS == 8) {
            float c8[8], c4[4], cf[2];
#pragma unroll
            for (int j = 0; j < 8; ++j) c8[j] = (h4 ? ca[j + 8] : ca[j]) + __shfl_xor_sync(0xffffffffu, h4 ? ca[j] : ca[j + 8], 16);
#pragma unroll
            for (int j = 0; j < 4; ++j) c4[j] = (h2 ? c8[j + 4] : c8[j]) + __shfl_xor_sync(0xffffffffu, h2 ? c8[j] : c8[j + 4], 8);
#pragma unroll
            for (int j = 0; j < 2; ++j) cf[j] = (h1 ? c4[j + 2] : c4[j]) + __shfl_xor_sync(0xffffffffu, h1 ? c4[j] : c4[j + 2], 4);
            *reinterpret_cast<float2*>(cb + 2 * lane) = make_float2(cf[0], cf[1]);       // columns 2 lane + {0, 1}
          } else {
            float c4[4], c2v[2];
#pragma unroll
            for (int j = 0; j < 4; ++j) c4[j] = (h4 ? ca[j + 4] : ca[j]) + __shfl_xor_sync(0xffffffffu, h4 ? ca[j] : ca[j + 4], 16);
#pragma unroll
            for (int j = 0; j < 2; ++j) c2v[j] = (h2 ? c4[j + 2] : c4[j]) + __shfl_xor_sync(0xffffffffu, h2 ? c4[j] : c4[j + 2], 8);
            const float cf = (h1 ? c2v[1] : c2v[0]) + __shfl_xor_sync(0xffffffffu, h1 ? c2v[0] : c2v[1], 4);
            // k = 2 * bit4 + bit3, b = bit2  ->  column 8 k + 2 t0 + b
            cb[8 * (2 * (int)h4 + (int)h2) + 2 * t0 + (int)h1] = cf;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_col(ncol % 3));
          pend_R = R; pend_J = J;
          ++ncol;
        }
        if (warp == 2 && lane == 0) SM3_TR(11, it);
      }
      // ---- row sums of this (CTA, row pair) segment: fold the NQ column groups in a fixed order ----
      const int kseg = (int)blockIdx.x - (int)(sym_flat_start(R, T, P) / p.sym_tpc);
      float rs[2][4];
#pragma unroll
      for (int rb = 0; rb < 2; ++rb) {
#pragma unroll
        for (int sl = 0; sl < 4; ++sl) {
          float lo, hi;
          f2_unpack(sum2[rb][sl], lo, hi);
          float v = lo + hi;
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          rs[rb][sl] = v;
          if (cg > 0 && t0 == 0) xsum[(cg - 1) * 256 + rb * 128 + q * 32 + t1 + 8 * (sl & 1) + 16 * (sl >> 1)] = v;
        }
      }
      named_bar_sync(1, 32 * NW);
      if (cg == 0 && t0 == 0) {
#pragma unroll
        for (int rb = 0; rb < 2; ++rb) {
#pragma unroll
          for (int sl = 0; sl < 4; ++sl) {
            const int rl = rb * 128 + q * 32 + t1 + 8 * (sl & 1) + 16 * (sl >> 1);
            float v = rs[rb][sl];
#pragma unroll
            for (int g = 1; g < NQ; ++g) v += xsum[(g - 1) * 256 + rl];
            p.partial[(int64_t)kseg * M + r0 + rl] = v;
          }
        }
      }
      named_bar_sync(1, 32 * NW);
      f += cnt;
    }
    fold_pending();
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) SM3_TR(7, 4);
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// symmetric forward ACROSS ranks (see MrPlan in common.cuh): the kernel above with (a) the per-row-pair tile sequence
// "own block (upper-triangular, local coordinates) | partner blocks | antipodal half block", (b) one flag wait per source
// rank in the TMA warp before the first tile of that rank's rows, (c) the column sums of partner tiles going to
// per-partner slabs (they are pushed to the owners by colsum_push_kernel).  Diagonal / positive masks only occur in the
// own block, in local row / column indices.
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ void mr_decode(const MrPlan& m, long f, int& R, int& off) {
  int lo = 0, hi = m.P_l - 1;                       // largest R with mr_prefix(R) <= f
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (mr_prefix(m, mid) <= f) lo = mid; else hi = mid - 1;
  }
  R = lo;
  off = (int)(f - mr_prefix(m, lo));
}
// t-th tile of row pair R: partner slot (-1 = own block) and the tile index in the owner's local tile order
__host__ __device__ __forceinline__ void mr_tile(const MrPlan& m, int R, int t, int& ps, int& j_l) {
  const int own = m.T_l - 2 * R;
  if (t < own) { ps = -1; j_l = 2 * R + t; return; }
  t -= own;
  const int q = t / m.T_l;
  if (q < m.H) { ps = q; j_l = t - q * m.T_l; return; }
  ps = m.H;
  j_l = t - m.H * m.T_l;                            // antipodal: [0, 2 (P_l / 2)) for anti == 1, [0, T_l) for anti == 2
}

template <int DP, int POLY, int NQ>
__global__ void __launch_bounds__(64 + 128 * NQ, 1)
infonce_tc_fwdsym_mr_kernel(const __grid_constant__ CUtensorMap tmap_cols, TcParams p) {
  using C = FwdCfg<DP>;
  constexpr int BN = C::BN, NSTAGE = C::NSTAGE;
  constexpr int D = 64 * DP;
  constexpr int NW = 4 * NQ;             // softmax warps
  constexpr int CW = 128 / NQ;           // S-tile columns per softmax warp
  constexpr int KS = CW / 8;             // 8-column groups per warp: registers 4 k .. 4 k + 3 of a 16x256b load
  const MrPlan& mr = p.mr;
  const int T = mr.T_l, P = mr.P_l;                 // LOCAL column tiles / row pairs
  const long f0 = (long)blockIdx.x * mr.tpc;
  const long f1 = min(mr.flat, f0 + mr.tpc);
  if (f0 >= f1) return;

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t sB = base;
  const uint32_t bars = sB + NSTAGE * C::STAGE;
  auto bar_full = [&](int i) { return bars + 8u * i; };
  auto bar_empty = [&](int i) { return bars + 8u * (NSTAGE + i); };
  auto bar_sfull = [&](int i) { return bars + 8u * (2 * NSTAGE + i); };
  auto bar_sempty = [&](int i) { return bars + 8u * (2 * NSTAGE + 2 + i); };
  const uint32_t bar_aready = bars + 8u * (2 * NSTAGE + 4);
  auto bar_col = [&](int i) { return bars + 8u * (2 * NSTAGE + 5 + i); };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + (bars - base) + 8u * (2 * NSTAGE + 8));
  float* xsum = reinterpret_cast<float*>(base_ptr + (bars - base) + 256);          // [NQ - 1][2 row blocks][128]
  float* cbuf = xsum + 768;                                                         // [3 ring slots][NW warps][CW]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned long long bar_limit = 2000000000ull + p.wait_timeout_ns;
  if (threadIdx.x == 0) SM3_TR(7, 0);

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(bar_full(i), 1); mbar_init(bar_empty(i), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(bar_sfull(i), 1); mbar_init(bar_sempty(i), NW); }
    mbar_init(bar_aready, NW);
    for (int i = 0; i < 3; ++i) mbar_init(bar_col(i), NW);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();
  pdl_launch();
  constexpr uint32_t kColA1 = 128, kColS = 256;

  if (warp == 0) {
    // =========================== TMA producer ===========================
    if (elect_one()) prefetch_tensormap(&tmap_cols);
    int it = 0;
    unsigned seen = 0u;                                          // partner slots whose rows are known to have landed
    for (long f = f0; f < f1;) {
      int R, off;
      mr_decode(mr, f, R, off);
      const int cnt = (int)min((long)(mr_count(mr, R) - off), f1 - f);
      for (int t = 0; t < cnt; ++t, ++it) {
        const int s = it % NSTAGE;
        const uint32_t ph = (uint32_t)(it / NSTAGE) & 1u;
        int ps, j_l;
        mr_tile(mr, R, off + t, ps, j_l);
        const int owner = ps < 0 ? mr.rank : mr_partner_rank(mr, ps);
        if (ps >= 0 && !((seen >> ps) & 1u)) {                   // first tile of this partner's rows: they must have landed
          if (elect_one()) {
            peer_flag_wait_one(p.wait_flags, p.wait_channel, owner, p.wait_epoch, p.wait_timeout_ns);
            fence_proxy_async_global();
          }
          __syncwarp();
          seen |= 1u << ps;
        }
        mbar_wait(bar_empty(s), ph ^ 1u, bar_limit);
        if (elect_one()) {
          mbar_expect_tx(bar_full(s), C::STAGE);
          const int row = mr_global_tile(mr, owner, j_l) * BN;
#pragma unroll
          for (int pnl = 0; pnl < DP; ++pnl)
            tma_load_2d(sB + s * C::STAGE + pnl * C::PANEL, &tmap_cols, bar_full(s), pnl * 64, row);
        }
        __syncwarp();
      }
      f += cnt;
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    constexpr uint32_t idesc = make_idesc_bf16(128, BN, 0, 0);
    constexpr uint32_t dhi = smem_desc_hi(1024);
    int it = 0, seg = 0;
    for (long f = f0; f < f1; ++seg) {
      int R, off;
      mr_decode(mr, f, R, off);
      const int cnt = (int)min((long)(mr_count(mr, R) - off), f1 - f);
      mbar_wait(bar_aready, (uint32_t)seg & 1u, bar_limit);      // this row pair's A operands are in TMEM
      tc_fence_after();
      for (int t = 0; t < cnt; ++t, ++it) {
        const int s = it % NSTAGE;
        const uint32_t sph = (uint32_t)it & 1u;
        mbar_wait(bar_full(s), (uint32_t)(it / NSTAGE) & 1u, bar_limit);
        if (it == 0) SM3_TR(7, 1);
        SM3_TR(0, it);
        const uint32_t lo0 = smem_desc_lo(sB + s * C::STAGE, 16);
#pragma unroll
        for (int rb = 0; rb < 2; ++rb) {
          mbar_wait(bar_sempty(rb), sph ^ 1u, bar_limit);
          tc_fence_after();
          SM3_TR(8 + rb, it);
          const uint32_t d_tmem = tmem + kColS + (uint32_t)rb * 128u;
          const uint32_t a_tmem = tmem + (uint32_t)rb * kColA1;
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < 4 * DP; ++ks)
              umma_ts(d_tmem, a_tmem + ks * 8, desc64(lo0 + (((ks >> 2) * C::PANEL + (ks & 3) * 32) >> 4), dhi), idesc,
                      ks > 0);
            if (rb == 1) umma_commit(bar_empty(s));
            umma_commit(bar_sfull(rb));
          }
          __syncwarp();
          SM3_TR(1 + rb, it);
        }
      }
      f += cnt;
    }
    SM3_TR(7, 2);
  } else {
    // =========================== softmax warps ===========================
    const int q = warp & 3;
    const int cg = (warp - 2) >> 2;                              // column group of this warp: columns [cg * CW, +CW)
    const int t0 = lane & 3, t1 = lane >> 2;
    const int row_in_tile = q * 32 + lane;                       // staging view: thread <-> TMEM lane
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint32_t lane_hi = (uint32_t)(q * 32 + 16) << 16;
    const float c2 = p.c2;
    const uint64_t c2p = f2_pack(c2, c2), nc2p = f2_pack(-c2, -c2);
    const bool clamp = 2.0f * c2 > 125.0f;
    const int64_t M = p.m_rows;                                  // local rows
    float* colpart = p.partial + (int64_t)mr.maxseg * M;         // own block: [P][M]; partner slot ps: [(1 + ps) P ...][M]
    int it = 0;
    // column sums are folded across the four lane quadrants one tile LATE (no CTA-wide rendezvous per tile): ring of 3
    // buffers, one mbarrier each (NW warp arrivals); pend_* describe the tile whose fold is outstanding
    int ncol = 0, pend_R = -1, pend_J = 0;                        // pend_R: slab row (own: R, partner ps: (1 + ps) P + R)
    auto fold_pending = [&]() {
      if (pend_R < 0) return;
      const int ci = (ncol - 1) % 3;
      mbar_wait(bar_col(ci), (uint32_t)((ncol - 1) / 3) & 1u, bar_limit);
      if (lane < 128 / NW) {                                     // NW warps x 128 / NW columns, quadrants in a fixed order
        const int c = (warp - 2) * (128 / NW) + lane;
        const float* src = cbuf + ci * 512 + (c / CW) * (4 * CW) + (c % CW);
        colpart[(int64_t)pend_R * M + (int64_t)pend_J * BN + c] = (src[0] + src[CW]) + (src[2 * CW] + src[3 * CW]);
      }
      pend_R = -1;
    };
    for (long f = f0; f < f1;) {
      int R, off;
      mr_decode(mr, f, R, off);
      const int cnt = (int)min((long)(mr_count(mr, R) - off), f1 - f);
      const int r0 = R * 256;
      // ---- stage this row pair into TMEM as A operands (all S-MMAs of the previous segment have completed: every
      //      softmax warp has waited for the last S tile) ----
#pragma unroll
      for (int rb = 0; rb < 2; ++rb) {
#pragma unroll
        for (int ch = 0; ch < DP; ++ch) {
          if ((ch % NQ) == cg) {
            uint32_t r[32];
            const uint4* src = reinterpret_cast<const uint4*>(p.z_rows + (size_t)(r0 + rb * 128 + row_in_tile) * D + ch * 64);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const uint4 v = __ldg(src + i);
              r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
            }
            tmem_st_x32(tmem + lane_addr + rb * kColA1 + ch * 32, r);
          }
        }
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_aready);

      // packed (column b = 0 | b = 1) running row sums of the thread's 4 rows per row block
      uint64_t sum2[2][4];
#pragma unroll
      for (int rb = 0; rb < 2; ++rb)
#pragma unroll
        for (int sl = 0; sl < 4; ++sl) sum2[rb][sl] = 0ull;
      for (int t = 0; t < cnt; ++t, ++it) {
        int ps, J;                                             // J: tile index in the column owner's LOCAL order
        mr_tile(mr, R, off + t, ps, J);
        const uint32_t sph = (uint32_t)it & 1u;
        const bool do_col = ps >= 0 || J > 2 * R + 1;
        uint64_t col2[KS];                                       // packed column sums: columns 8 k + 2 t0 + {0, 1}
        // TMEM loads run one half tile ahead of the exponentials: vb(rb) is in flight during the first half of row block
        // rb, and va of row block 1 (its S tile has been ready for a while: the MMA warp runs ahead) during the second
        // half of row block 0 -- only the first load of a tile is exposed.
        uint32_t va[4 * KS], vb[4 * KS];
        auto begin_rb = [&](int rb) {
          mbar_wait(bar_sfull(rb), sph, bar_limit);
          tc_fence_after();
          if (warp == 2 && lane == 0) SM3_TR(3 + 2 * rb, it);
          tmem_ld_16x256b(tmem + kColS + (uint32_t)rb * 128u + (uint32_t)(cg * CW) + lane_addr, va);   // rows t1, t1 + 8
        };
        begin_rb(0);
#pragma unroll
        for (int rb = 0; rb < 2; ++rb) {
          const uint32_t taddr = tmem + kColS + (uint32_t)rb * 128u + (uint32_t)(cg * CW);
          tmem_ld_wait(va);
          tmem_ld_16x256b(taddr + lane_hi, vb);                  // rows t1 + 16, t1 + 24: in flight during the first half
          int Jp = 2 * R + rb + P;                               // the tile that holds these rows' positives
          if (Jp >= T) Jp -= T;
          const bool special = ps < 0 && ((J == 2 * R + rb) || (J == Jp));
#pragma unroll
          for (int hv = 0; hv < 2; ++hv) {
            if (hv == 1) {
              tmem_ld_wait(vb);
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar_sempty(rb));        // S stage free (the MMA warp runs a tile ahead anyway)
              if (warp == 2 && lane == 0) SM3_TR(4 + 2 * rb, it);
              if (rb == 0) begin_rb(1);                          // va is dead: prefetch row block 1's first half
            }
            const uint32_t* v = hv == 0 ? va : vb;
            if (!special) {
              uint64_t e[2 * KS];
#pragma unroll
              for (int k = 0; k < KS; ++k) {
                e[2 * k] = f2_fma(f2_pack_u(v[4 * k], v[4 * k + 1]), c2p, nc2p);
                e[2 * k + 1] = f2_fma(f2_pack_u(v[4 * k + 2], v[4 * k + 3]), c2p, nc2p);
              }
              // POLY of every 8 exponentials on the FMA pipes: whole pairs (4: every other pair; 2: every fourth pair)
#pragma unroll
              for (int i = 0; i < 2 * KS; ++i)
                e[i] = ((POLY >= 4 && (i & 1) == 0) || (POLY >= 2 && POLY < 4 && (i & 3) == 0)) ? f2_ex2_fma(e[i], clamp) : f2_ex2(e[i]);
#pragma unroll
              for (int k = 0; k < KS; ++k) {
                sum2[rb][2 * hv] = f2_add(sum2[rb][2 * hv], e[2 * k]);
                sum2[rb][2 * hv + 1] = f2_add(sum2[rb][2 * hv + 1], e[2 * k + 1]);
                const uint64_t cs = f2_add(e[2 * k], e[2 * k + 1]);
                col2[k] = (rb == 0 && hv == 0) ? cs : f2_add(col2[k], cs);
              }
            } else {
              // the tile holds the diagonal or the positives of these rows: mask per element
              const int rbase = r0 + rb * 128 + q * 32 + t1 + 16 * hv;
              const int cbase = J * BN + cg * CW + 2 * t0;
#pragma unroll
              for (int k = 0; k < KS; ++k) {
                uint64_t cs = 0ull;
#pragma unroll
                for (int a8 = 0; a8 < 2; ++a8) {
                  const int row = rbase + 8 * a8;
                  const int pj = positive_of(row, p.n_local);    // local coordinates: the own block is an n_local problem
                  float ev[2];
#pragma unroll
                  for (int bb = 0; bb < 2; ++bb) {
                    const int col = cbase + 8 * k + bb;
                    const float sv = __uint_as_float(v[4 * k + 2 * a8 + bb]);
                    const bool is_pos = (col == pj);
                    if (is_pos && pj > row) {                    // S is symmetric: one read serves both rows of the pair
                      const float pv = sv * p.inv_T;
                      p.pos[row] = pv;
                      p.pos[pj] = pv;
                    }
                    ev[bb] = (is_pos || col == row) ? 0.f : ex2(fmaf(sv, c2, -c2));
                  }
                  const uint64_t e = f2_pack(ev[0], ev[1]);
                  sum2[rb][2 * hv + a8] = f2_add(sum2[rb][2 * hv + a8], e);
                  cs = f2_add(cs, e);
                }
                col2[k] = (rb == 0 && hv == 0) ? cs : f2_add(col2[k], cs);
              }
            }
          }
        }
        if (warp == 2 && lane == 0) SM3_TR(10, it);
        fold_pending();                                          // the previous column-sum tile: every warp arrived long ago
        if (do_col) {
          // column sums over this warp's 32 rows x 2 row blocks: butterfly over the lanes t1 = lane / 4 that share a
          // column (lane bit 4 <-> k bit 2 or 1, ...); each step halves the values a lane carries
          float ca[2 * KS];
#pragma unroll
          for (int k = 0; k < KS; ++k) f2_unpack(col2[k], ca[2 * k], ca[2 * k + 1]);
          const bool h4 = (lane & 16) != 0, h2 = (lane & 8) != 0, h1 = (lane & 4) != 0;
          float* cb = cbuf + (ncol % 3) * 512 + (cg * 4 + q) * CW;
          if constexpr (KS == 8) {
            float c8[8], c4[4], cf[2];
#pragma unroll
            for (int j = 0; j < 8; ++j) c8[j] = (h4 ? ca[j + 8] : ca[j]) + __shfl_xor_sync(0xffffffffu, h4 ? ca[j] : ca[j + 8], 16);
#pragma unroll
            for (int j = 0; j < 4; ++j) c4[j] = (h2 ? c8[j + 4] : c8[j]) + __shfl_xor_sync(0xffffffffu, h2 ? c8[j] : c8[j + 4], 8);
#pragma unroll
            for (int j = 0; j < 2; ++j) cf[j] = (h1 ? c4[j + 2] : c4[j]) + __shfl_xor_sync(0xffffffffu, h1 ? c4[j] : c4[j + 2], 4);
            *reinterpret_cast<float2*>(cb + 2 * lane) = make_float2(cf[0], cf[1]);       // columns 2 lane + {0, 1}
          } else {
            float c4[4], c2v[2];
#pragma unroll
            for (int j = 0; j < 4; ++j) c4[j] = (h4 ? ca[j + 4] : ca[j]) + __shfl_xor_sync(0xffffffffu, h4 ? ca[j] : ca[j + 4], 16);
#pragma unroll
            for (int j = 0; j < 2; ++j) c2v[j] = (h2 ? c4[j + 2] : c4[j]) + __shfl_xor_sync(0xffffffffu, h2 ? c4[j] : c4[j + 2], 8);
            const float cf = (h1 ? c2v[1] : c2v[0]) + __shfl_xor_sync(0xffffffffu, h1 ? c2v[0] : c2v[1], 4);
            // k = 2 * bit4 + bit3, b = bit2  ->  column 8 k + 2 t0 + b
            cb[8 * (2 * (int)h4 + (int)h2) + 2 * t0 + (int)h1] = cf;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_col(ncol % 3));
          pend_R = (ps < 0 ? 0 : (1 + ps) * P) + R; pend_J = J;
          ++ncol;
        }
        if (warp == 2 && lane == 0) SM3_TR(11, it);
      }
      // ---- row sums of this (CTA, row pair) segment: fold the NQ column groups in a fixed order ----
      const int kseg = (int)blockIdx.x - (int)(mr_prefix(mr, R) / mr.tpc);
      float rs[2][4];
#pragma unroll
      for (int rb = 0; rb < 2; ++rb) {
#pragma unroll
        for (int sl = 0; sl < 4; ++sl) {
          float lo, hi;
          f2_unpack(sum2[rb][sl], lo, hi);
          float v = lo + hi;
          v += __shfl_xor_sync(0xffffffffu, v, 1);
          v += __shfl_xor_sync(0xffffffffu, v, 2);
          rs[rb][sl] = v;
          if (cg > 0 && t0 == 0) xsum[(cg - 1) * 256 + rb * 128 + q * 32 + t1 + 8 * (sl & 1) + 16 * (sl >> 1)] = v;
        }
      }
      named_bar_sync(1, 32 * NW);
      if (cg == 0 && t0 == 0) {
#pragma unroll
        for (int rb = 0; rb < 2; ++rb) {
#pragma unroll
          for (int sl = 0; sl < 4; ++sl) {
            const int rl = rb * 128 + q * 32 + t1 + 8 * (sl & 1) + 16 * (sl >> 1);
            float v = rs[rb][sl];
#pragma unroll
            for (int g = 1; g < NQ; ++g) v += xsum[(g - 1) * 256 + rl];
            p.partial[(int64_t)kseg * M + r0 + rl] = v;
          }
        }
      }
      named_bar_sync(1, 32 * NW);
      f += cnt;
    }
    fold_pending();
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) SM3_TR(7, 4);
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
// NS = S/H stages in TMEM.  2 at D = 256 (all 512 columns in use).  At D <= 128 TMEM has room for 4, which keeps four
// tiles of the chain  S-MMA -> tcgen05.ld -> exp -> tcgen05.st -> dZ-MMA  in flight.  Measured on B200 at cfg2
// (4096 x 128): 83.7 -> 81.5 us for the backward stage, i.e. the chain latency is NOT what holds this shape at 0.7 us
// per 128 x 64 tile (MMA work: 0.27 us); the next suspect is the issue rate of the small N = 64 / K = 16 MMAs.
template <int DP, int NS> struct BwdCfg {
  static constexpr int BN = 64;
  static constexpr uint32_t PANEL = BN * 128;          // 8 KB: 64 rows x 64 bf16
  static constexpr uint32_t STAGE = DP * PANEL;        // <= 32 KB
  static constexpr int NSTAGE = NS + 2;                // smem ring: tiles it .. it+NS are live, one more in flight
  static constexpr uint32_t SMEM = NSTAGE * STAGE + 1024 + 256;
  // TMEM columns: A [0, 32*DP) | S/H stage i [128 + 64 i, 192 + 64 i), i < NS | dZ [128 + 64 NS, 128 + 64 NS + 64*DP)
  static constexpr uint32_t kColS = 128, kColDZ = 128 + 64 * NS;
  static_assert(kColDZ + 64 * DP <= 512, "TMEM budget");
  static_assert(8 * (2 * NSTAGE + 2 * NS + 3) <= 256, "barrier block");
};

__global__ void tc_bwd_prep_kernel(const float* __restrict__ glse, const float* __restrict__ nsum, int stride,
                                   int m_cols, int m_pad, float* __restrict__ acol) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m_pad) return;
  float a = 0.f;
  if (j < m_cols) { const float s = nsum[(size_t)j * stride]; a = s > 0.f ? glse[(size_t)j * stride] / s : 0.f; }
  acol[j] = a;
}

// kWait (fused exchange): the per-column statistics (acol, gpos_c) are written by their owner ranks over NVLink while
// this kernel is already running; every softmax warp waits for the peers' flags before the first column tile another
// rank owns and reads the statistics with L2-coherent loads (ld.global.cg) instead of the read-only path.
template <int DP, int NG, bool kWait, int NS>
__global__ void __launch_bounds__(64 + 256 * NG, 1)
infonce_tc_bwd_kernel(const __grid_constant__ CUtensorMap tmap_cols, TcParams p) {
  using C = BwdCfg<DP, NS>;
  static_assert(NS % NG == 0, "each softmax group must keep its own S/H stages");
  constexpr int BN = C::BN, NSTAGE = C::NSTAGE;
  constexpr int D = 64 * DP;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t sB = base;
  const uint32_t bars = sB + NSTAGE * C::STAGE;
  auto bar_full = [&](int i) { return bars + 8u * i; };
  auto bar_empty = [&](int i) { return bars + 8u * (NSTAGE + i); };
  auto bar_sfull = [&](int i) { return bars + 8u * (2 * NSTAGE + i); };
  auto bar_hfull = [&](int i) { return bars + 8u * (2 * NSTAGE + NS + i); };
  const uint32_t bar_aready = bars + 8u * (2 * NSTAGE + 2 * NS);
  const uint32_t bar_dzfull = bars + 8u * (2 * NSTAGE + 2 * NS + 1);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + (bars - base) + 8u * (2 * NSTAGE + 2 * NS + 2));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned long long bar_limit = 2000000000ull + p.wait_timeout_ns;   // + cross-rank bound in fused-exchange mode
  const int r0 = blockIdx.x * kBM;
  const int split = blockIdx.y;
  const int t_begin = split * p.tiles_per_split;
  const int t_end = min(p.col_tiles, t_begin + p.tiles_per_split);
  const int n_tiles = t_end - t_begin;
  // Column tiles are visited in a per-CTA rotated order.  All row blocks of a wave would otherwise sweep the SAME
  // 32-64 KB tile at the same time and serialise on the few L2 slices holding it (measured: the forward kernel was
  // bimodal, 1.7 ms or 3.5-6 ms at cfg4, depending on whether the CTAs happened to run in lockstep).
  const int rot = tile_rotation(p, t_begin, n_tiles);
  auto tile_of = [&](int it) {
    int t = it + rot;
    t = t_begin + (t >= n_tiles ? t - n_tiles : t);
    if (t >= p.skip_a) t += p.skip_len;        // skip mode: hop over the column tiles owned by this rank
    if (t >= p.skip_b) t += p.skip_len;
    return t;
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(bar_full(i), 1); mbar_init(bar_empty(i), 1); }
    for (int i = 0; i < NS; ++i) { mbar_init(bar_sfull(i), 1); mbar_init(bar_hfull(i), 8); }
    mbar_init(bar_aready, 8 * NG);
    mbar_init(bar_dzfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();        // PDL: everything above overlapped the preceding kernel's tail; global memory from here on
  pdl_launch();
  constexpr uint32_t kColS = C::kColS, kColDZ = C::kColDZ;

  if (warp == 0) {
    // =========================== TMA producer (warp-uniform loop, one elected lane issues) ===========
    if (elect_one()) prefetch_tensormap(&tmap_cols);
    for (int it = 0; it < n_tiles; ++it) {
      const int s = it % NSTAGE;
      const uint32_t ph = (uint32_t)(it / NSTAGE) & 1u;
      mbar_wait(bar_empty(s), ph ^ 1u, bar_limit);
      if (elect_one()) {
        mbar_expect_tx(bar_full(s), C::STAGE);
        const int row = tile_of(it) * BN;
#pragma unroll
        for (int pnl = 0; pnl < DP; ++pnl)
          tma_load_2d(sB + s * C::STAGE + pnl * C::PANEL, &tmap_cols, bar_full(s), pnl * 64, row);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (warp-uniform loop, one elected lane issues) ============
    constexpr uint32_t idesc_s = make_idesc_bf16(128, BN, 0, 0);     // S  = Zr  * Zc^T   (B K-major)
    constexpr uint32_t idesc_z = make_idesc_bf16(128, D, 0, 1);      // dZ += H  * Zc     (B MN-major)
    constexpr uint32_t dhi = smem_desc_hi(1024);
    auto issue_s = [&](int it) {
      const int s = it % NSTAGE, as = it % NS;
      mbar_wait(bar_full(s), (uint32_t)(it / NSTAGE) & 1u, bar_limit);
      tc_fence_after();
      const uint32_t d_tmem = tmem + kColS + (uint32_t)as * 64u;
      const uint32_t lo0 = smem_desc_lo(sB + s * C::STAGE, 16);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4 * DP; ++ks)
          umma_ts(d_tmem, tmem + ks * 8, desc64(lo0 + (((ks >> 2) * C::PANEL + (ks & 3) * 32) >> 4), dhi), idesc_s,
                  ks > 0);
        umma_commit(bar_sfull(as));
      }
      __syncwarp();
    };
    mbar_wait(bar_aready, 0, bar_limit);
    tc_fence_after();
#pragma unroll
    for (int i = 0; i < NS; ++i)
      if (i < n_tiles) issue_s(i);
    for (int it = 0; it < n_tiles; ++it) {
      const int s = it % NSTAGE, as = it % NS;
      mbar_wait(bar_hfull(as), (uint32_t)(it / NS) & 1u, bar_limit);
      tc_fence_after();
      SM3_TR(0, it);
      const uint32_t h_tmem = tmem + kColS + (uint32_t)as * 64u;
      const uint32_t lo0 = smem_desc_lo(sB + s * C::STAGE, C::PANEL);   // MN-major: LBO = panel stride
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < BN / 16; ++ks) {
          // H k-columns [0,32) live at TMEM cols +0..15, k-columns [32,64) at +32..47 (written by the two warp halves)
          const uint32_t a_tmem = h_tmem + (ks < 2 ? ks * 8 : 32 + (ks - 2) * 8);
          umma_ts(tmem + kColDZ, a_tmem, desc64(lo0 + ks * (2048 >> 4), dhi), idesc_z, (it > 0 || ks > 0) ? 1u : 0u);
        }
        umma_commit(bar_empty(s));
      }
      __syncwarp();
      SM3_TR(1, it);
      if (it + NS < n_tiles) issue_s(it + NS);   // in-order tensor pipe: overwrites S/H stage `as` only after dZ(it)
      SM3_TR(2, it);
    }
    if (elect_one()) umma_commit(bar_dzfull);
    __syncwarp();
  } else {
    // =========================== softmax / epilogue warps ===========================
    const int q = warp & 3;
    const int grp = (warp - 2) >> 3;           // softmax group: owns tiles with it % NG == grp
    const int half = ((warp - 2) >> 2) & 1;    // which 32 of the tile's 64 columns
    const int combo = grp * 2 + half;
    const int row_in_tile = q * 32 + lane;
    const int l = r0 + row_in_tile;
    const bool valid = l < p.m_rows;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;

#pragma unroll
    for (int ch = 0; ch < DP; ++ch) {
      if ((ch % (2 * NG)) == combo) {
        uint32_t r[32];
        if (valid) {
          const uint4* src = reinterpret_cast<const uint4*>(p.z_rows + (size_t)l * D + ch * 64);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint4 v = __ldg(src + i);
            r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = 0u;
        }
        tmem_st_x32(tmem + lane_addr + ch * 32, r);
      }
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_aready);

    const int g = valid ? global_row(l, p.n_local, p.pair_offset, p.n_global) : -1;
    const int pj = valid ? positive_of(g, p.n_global) : -1;
    float a_i = 0.f, gp_i = 0.f;
    if (valid) {
      const float s = p.nsum_r[l];
      a_i = s > 0.f ? p.glse_r[l] / s : 0.f;
      gp_i = p.gpos_r[l];
    }
    const float c2 = p.c2;

    bool remote_ready = !kWait;
    for (int it = grp; it < n_tiles; it += NG) {
      const int as = it % NS;
      const int tile = tile_of(it);
      if constexpr (kWait) {
        if (!remote_ready && !tile_is_local(p, tile)) {   // the owners' statistics must have landed
          if (lane == 0) peer_flags_wait_all(p.wait_flags, p.wait_world, p.wait_channel, p.wait_epoch, p.wait_timeout_ns);
          __syncwarp();
          remote_ready = true;
        }
      }
      mbar_wait(bar_sfull(as), (uint32_t)(it / NS) & 1u, bar_limit);
      tc_fence_after();
      if (warp == 2 && lane == 0) SM3_TR(3, it);
      const uint32_t taddr = tmem + lane_addr + kColS + (uint32_t)as * 64u + (uint32_t)half * 32u;
      uint32_t v[32];
      tmem_ld_x32(taddr, v);
      tmem_ld_wait(v);
      if (warp == 2 && lane == 0) SM3_TR(4, it);
      const int cb = tile * BN + half * 32;
      const bool need = (cb + 32 > p.m_cols) ||
                        (valid && ((unsigned)(g - cb) < 32u || (unsigned)(pj - cb) < 32u));
      uint32_t h[16];
      const float4* ap = reinterpret_cast<const float4*>(p.acol + cb);
      if (!__any_sync(0xffffffffu, need)) {
        // packed fp32 (FFMA2 / FADD2 / FMUL2 on register pairs): 13 issue slots per four logits instead of 19 -- the
        // kernel is bound by its MMA chain, not by these warps, but it runs at the power cap
        const uint64_t c2p = f2_pack(c2, c2), nc2p = f2_pack(-c2, -c2), aip = f2_pack(a_i, a_i);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 aj = kWait ? __ldcg(ap + i) : __ldg(ap + i);
          const uint64_t e01 = f2_ex2(f2_fma(f2_pack_u(v[4 * i], v[4 * i + 1]), c2p, nc2p));
          const uint64_t e23 = f2_ex2(f2_fma(f2_pack_u(v[4 * i + 2], v[4 * i + 3]), c2p, nc2p));
          float h0, h1, h2, h3;
          f2_unpack(f2_mul(e01, f2_add(aip, f2_pack(aj.x, aj.y))), h0, h1);
          f2_unpack(f2_mul(e23, f2_add(aip, f2_pack(aj.z, aj.w))), h2, h3);
          h[2 * i] = pack_bf16x2(h0, h1);
          h[2 * i + 1] = pack_bf16x2(h2, h3);
        }
      } else {
        float hv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int col = cb + i;
          const float s = __uint_as_float(v[i]);
          float x = 0.f;
          if (valid && col < p.m_cols && col != g) {
            if (col == pj) x = gp_i + (kWait ? __ldcg(p.gpos_c + (size_t)col * p.cstride) : p.gpos_c[(size_t)col * p.cstride]);
            else x = ex2(fmaf(s, c2, -c2)) * (a_i + (kWait ? __ldcg(p.acol + col) : p.acol[col]));
          }
          hv[i] = x;
        }
#pragma unroll
        for (int i = 0; i < 16; ++i) h[i] = pack_bf16x2(hv[2 * i], hv[2 * i + 1]);
      }
      if (warp == 2 && lane == 0) SM3_TR(5, it);
      tmem_st_x16(taddr, h);          // H overwrites this thread's own (already consumed) S columns
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_hfull(as));
      if (warp == 2 && lane == 0) SM3_TR(6, it);
    }

    // ---- epilogue: dZ (TMEM fp32) * 1/T -> global partial ----
    mbar_wait(bar_dzfull, 0, bar_limit);
    tc_fence_after();
    float* dst = p.dz_partial + ((size_t)split * p.m_rows + (valid ? l : 0)) * D;
#pragma unroll
    for (int ch = 0; ch < 2 * DP; ++ch) {
      if ((ch % (2 * NG)) == combo) {
        uint32_t v[32];
        tmem_ld_x32(tmem + lane_addr + kColDZ + ch * 32, v);
        tmem_ld_wait(v);
        if (valid) {
          float4* o = reinterpret_cast<float4*>(dst + ch * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            o[i] = make_float4(__uint_as_float(v[4 * i]) * p.inv_T, __uint_as_float(v[4 * i + 1]) * p.inv_T,
                               __uint_as_float(v[4 * i + 2]) * p.inv_T, __uint_as_float(v[4 * i + 3]) * p.inv_T);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// backward, second form: TILE-ALTERNATING softmax groups + per-column statistics through shared memory.
//
// What held the first form at ~1600 clk per 128 x 64 tile at D = 128 (MMA work: 512, MUFU work: 512): all eight softmax
// warps walked every tile together (the two warps of an SM sub-partition split the tile's columns), so per tile each
// sub-partition paid, back to back and with nothing to overlap them,  tcgen05.ld latency -> an L2 round trip for the 64
// a_j values (LDG issued after the barrier) -> 512 MUFU cycles -> tcgen05.st latency.  Here
//   * warps 2-5 own the even tiles and warps 6-9 the odd ones (one thread = one row x all 64 columns of the tile), so
//     the two warps sharing a sub-partition are always in DIFFERENT phases of DIFFERENT tiles and the MUFU pipe
//     stays busy while the other warp sits in its TMEM round trips;
//   * a_j travels with the column tile: the TMA warp adds one 256-byte bulk copy per stage and the softmax threads
//     read it with broadcast LDS.128 -- no global-memory latency inside the softmax loop at all.  (Fused exchange: the
//     TMA warp, not the softmax warps, waits for the owners' statistics before the first remote tile.)
// TMEM: A [0, 32 DP) | S/H stage i [128 + 64 i, 192 + 64 i) | dZ [128 + 64 NS, 128 + 64 NS + 64 DP).  H (bf16 pairs)
// overwrites the first 32 columns of the S stage it was computed from.
// ------------------------------------------------------------------------------------------------
template <int DP, int NS, int BN_> struct Bwd2Cfg {
  static constexpr int BN = BN_;                       // columns per tile: 64, or 128 (D <= 128: TMEM has the room)
  static constexpr uint32_t PANEL = BN * 128;          // BN rows x 64 bf16
  static constexpr uint32_t STAGE = DP * PANEL;        // <= 32 KB
  static constexpr uint32_t ACOL = BN * 4;             // a_j of the tile's columns
  static constexpr int NSTAGE = NS + 2;                // smem ring: tiles it .. it+NS are live, one more in flight
  static constexpr uint32_t SMEM = NSTAGE * (STAGE + ACOL) + 1024 + 256;
  static constexpr uint32_t kColS = 128, kColDZ = 128 + BN * NS;
  static_assert(BN == 64 || BN == 128, "tile width");
  static_assert(32 * DP <= 128, "A operand region");
  static_assert(kColDZ + 64 * DP <= 512, "TMEM budget");
  static_assert(8 * (2 * NSTAGE + 2 * NS + 3) <= 256, "barrier block");
};

// COLSPLIT (needs BN = 128): every softmax warp works on EVERY tile and the two groups split its columns (group g owns
// columns [64 g, 64 g + 64)), so one tile's exponentials occupy all four MUFU pipes while the tensor pipe runs the other
// stage's dZ / S MMAs -- with only two 128-column stages in TMEM, tile-alternating groups cannot hide a tile's
// ~2000-cycle softmax latency (measured: 1935 cycles per tile), the column split halves it.
template <int DP, bool kWait, int NS, int BNT, int POLY, bool COLSPLIT>
__global__ void __launch_bounds__(320, 1)
infonce_tc_bwd2_kernel(const __grid_constant__ CUtensorMap tmap_cols, TcParams p) {
  using C = Bwd2Cfg<DP, NS, BNT>;
  static_assert(COLSPLIT || NS % 2 == 0, "each softmax group keeps its own S/H stages");
  static_assert(!COLSPLIT || BNT == 128, "the column split is between two 64-column halves");
  constexpr int BN = C::BN, NSTAGE = C::NSTAGE;
  constexpr int D = 64 * DP;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* base_ptr = smem_raw + (base - raw);
  const uint32_t sB = base;
  const uint32_t sA = sB + NSTAGE * C::STAGE;                       // a_j slots, 256 B per stage
  const uint32_t bars = sA + NSTAGE * C::ACOL;
  auto bar_full = [&](int i) { return bars + 8u * i; };
  auto bar_empty = [&](int i) { return bars + 8u * (NSTAGE + i); };
  auto bar_sfull = [&](int i) { return bars + 8u * (2 * NSTAGE + i); };
  auto bar_hfull = [&](int i) { return bars + 8u * (2 * NSTAGE + NS + i); };
  const uint32_t bar_aready = bars + 8u * (2 * NSTAGE + 2 * NS);
  const uint32_t bar_dzfull = bars + 8u * (2 * NSTAGE + 2 * NS + 1);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(base_ptr + (bars - base) + 8u * (2 * NSTAGE + 2 * NS + 2));
  const float* acol_smem = reinterpret_cast<const float*>(base_ptr + (sA - base));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) SM3_TR(7, 0);                                      // trace: CTA start
  const unsigned long long bar_limit = 2000000000ull + p.wait_timeout_ns;   // + cross-rank bound in fused-exchange mode
  const int r0 = blockIdx.x * kBM;
  const int split = blockIdx.y;
  const int t_begin = split * p.tiles_per_split;
  const int t_end = min(p.col_tiles, t_begin + p.tiles_per_split);
  const int n_tiles = t_end - t_begin;
  const int rot = tile_rotation(p, t_begin, n_tiles);   // decorrelates the CTAs' L2 accesses; local tiles first (fused)
  auto tile_of = [&](int it) {
    int t = it + rot;
    t = t_begin + (t >= n_tiles ? t - n_tiles : t);
    if (t >= p.skip_a) t += p.skip_len;        // skip mode: hop over the column tiles owned by this rank
    if (t >= p.skip_b) t += p.skip_len;
    return t;
  };

  if (threadIdx.x == 0) {
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(bar_full(i), 1); mbar_init(bar_empty(i), 1); }
    for (int i = 0; i < NS; ++i) { mbar_init(bar_sfull(i), 1); mbar_init(bar_hfull(i), COLSPLIT ? 8 : 4); }
    mbar_init(bar_aready, 8);
    mbar_init(bar_dzfull, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();        // PDL: everything above overlapped the preceding kernel's tail; global memory from here on
  pdl_launch();
  constexpr uint32_t kColS = C::kColS, kColDZ = C::kColDZ;

  if (warp == 0) {
    // =========================== TMA producer: column tile + its 64 a_j ===========================
    if (elect_one()) prefetch_tensormap(&tmap_cols);
    bool remote_ready = !kWait;
    for (int it = 0; it < n_tiles; ++it) {
      const int s = it % NSTAGE;
      const uint32_t ph = (uint32_t)(it / NSTAGE) & 1u;
      const int tile = tile_of(it);
      if constexpr (kWait) {
        if (!remote_ready && !tile_is_local(p, tile)) {   // the owners' statistics (a_j) must have landed
          if (elect_one()) {
            peer_flags_wait_all(p.wait_flags, p.wait_world, p.wait_channel, p.wait_epoch, p.wait_timeout_ns);
            fence_proxy_async_global();
          }
          __syncwarp();
          remote_ready = true;
        }
      }
      mbar_wait(bar_empty(s), ph ^ 1u, bar_limit);
      if (elect_one()) {
        mbar_expect_tx(bar_full(s), C::STAGE + C::ACOL);
        const int row = tile * BN;
#pragma unroll
        for (int pnl = 0; pnl < DP; ++pnl)
          tma_load_2d(sB + s * C::STAGE + pnl * C::PANEL, &tmap_cols, bar_full(s), pnl * 64, row);
        bulk_load_1d(sA + s * C::ACOL, p.acol + row, C::ACOL, bar_full(s));
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ===========================
    constexpr uint32_t idesc_s = make_idesc_bf16(128, BN, 0, 0);     // S  = Zr  * Zc^T   (B K-major)
    constexpr uint32_t idesc_z = make_idesc_bf16(128, D, 0, 1);      // dZ += H  * Zc     (B MN-major)
    constexpr uint32_t dhi = smem_desc_hi(1024);
    auto issue_s = [&](int it) {
      const int s = it % NSTAGE, as = it % NS;
      mbar_wait(bar_full(s), (uint32_t)(it / NSTAGE) & 1u, bar_limit);
      tc_fence_after();
      const uint32_t d_tmem = tmem + kColS + (uint32_t)as * (uint32_t)BN;
      const uint32_t lo0 = smem_desc_lo(sB + s * C::STAGE, 16);
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < 4 * DP; ++ks)
          umma_ts(d_tmem, tmem + ks * 8, desc64(lo0 + (((ks >> 2) * C::PANEL + (ks & 3) * 32) >> 4), dhi), idesc_s,
                  ks > 0);
        umma_commit(bar_sfull(as));
      }
      __syncwarp();
    };
    mbar_wait(bar_aready, 0, bar_limit);
    tc_fence_after();
    SM3_TR(7, 1);                                                          // trace: rows staged, first MMA
#pragma unroll
    for (int i = 0; i < NS; ++i)
      if (i < n_tiles) issue_s(i);
    for (int it = 0; it < n_tiles; ++it) {
      const int s = it % NSTAGE, as = it % NS;
      mbar_wait(bar_hfull(as), (uint32_t)(it / NS) & 1u, bar_limit);
      tc_fence_after();
      SM3_TR(0, it);
      const uint32_t h_tmem = tmem + kColS + (uint32_t)as * (uint32_t)BN;   // BN k-values packed in BN / 2 columns
      const uint32_t lo0 = smem_desc_lo(sB + s * C::STAGE, C::PANEL);   // MN-major: LBO = panel stride
      if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < BN / 16; ++ks)
          umma_ts(tmem + kColDZ, h_tmem + ks * 8, desc64(lo0 + ks * (2048 >> 4), dhi), idesc_z,
                  (it > 0 || ks > 0) ? 1u : 0u);
        umma_commit(bar_empty(s));
      }
      __syncwarp();
      SM3_TR(1, it);
      if (it + NS < n_tiles) issue_s(it + NS);   // in-order tensor pipe: overwrites S/H stage `as` only after dZ(it)
      SM3_TR(2, it);
    }
    if (elect_one()) umma_commit(bar_dzfull);
    __syncwarp();
    SM3_TR(7, 2);                                                          // trace: last MMA issued
  } else {
    // =========================== softmax / epilogue warps ===========================
    const int q = warp & 3;                    // TMEM lane quadrant
    const int grp = (warp - 2) >> 2;           // group 0 = warps 2-5 (even tiles), group 1 = warps 6-9 (odd tiles)
    const int row_in_tile = q * 32 + lane;
    const int l = r0 + row_in_tile;
    const bool valid = l < p.m_rows;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;

#pragma unroll
    for (int ch = 0; ch < DP; ++ch) {
      if ((ch & 1) == grp) {
        uint32_t r[32];
        if (valid) {
          const uint4* src = reinterpret_cast<const uint4*>(p.z_rows + (size_t)l * D + ch * 64);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const uint4 v = __ldg(src + i);
            r[4 * i] = v.x; r[4 * i + 1] = v.y; r[4 * i + 2] = v.z; r[4 * i + 3] = v.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = 0u;
        }
        tmem_st_x32(tmem + lane_addr + ch * 32, r);
      }
    }
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_aready);

    const int g = valid ? global_row(l, p.n_local, p.pair_offset, p.n_global) : -1;
    const int pj = valid ? positive_of(g, p.n_global) : -1;
    float a_i = 0.f, gp_i = 0.f;
    if (valid) {
      const float s = p.nsum_r[l];
      a_i = s > 0.f ? p.glse_r[l] / s : 0.f;
      gp_i = p.gpos_r[l];
    }
    const float c2 = p.c2;

    for (int it = COLSPLIT ? 0 : grp; it < n_tiles; it += COLSPLIT ? 1 : 2) {
      const int as = it % NS, s = it % NSTAGE;
      const int tile = tile_of(it);
      mbar_wait(bar_sfull(as), (uint32_t)(it / NS) & 1u, bar_limit);
      mbar_wait(bar_full(s), (uint32_t)(it / NSTAGE) & 1u, bar_limit);   // completed long ago: makes the TMA-written a_j ours
      tc_fence_after();
      if ((warp == 2 || (!COLSPLIT && warp == 6)) && lane == 0) SM3_TR(3, it);
      const uint32_t tbase = tmem + lane_addr + kColS + (uint32_t)as * (uint32_t)BN;
#pragma unroll
      for (int hc0 = 0; hc0 < (COLSPLIT ? 1 : BN / 64); ++hc0) {   // 64 columns at a time
        const int hc = COLSPLIT ? grp : hc0;
        const uint32_t taddr = tbase + (uint32_t)hc * 64u;
        uint32_t v0[32], v1[32];
        tmem_ld_x32(taddr, v0);
        tmem_ld_x32(taddr + 32, v1);
        tmem_ld_wait(v0);
        tmem_ld_wait(v1);
        if (hc0 == 0 && (warp == 2 || (!COLSPLIT && warp == 6)) && lane == 0) SM3_TR(4, it);
        const int cb = tile * BN + hc * 64;
        const bool need = (cb + 64 > p.m_cols) ||
                          (valid && ((unsigned)(g - cb) < 64u || (unsigned)(pj - cb) < 64u));
        const float* as_ = acol_smem + s * BN + hc * 64;
        const float4* ap = reinterpret_cast<const float4*>(as_);
        uint32_t h[32];
        if (!__any_sync(0xffffffffu, need)) {
          // packed fp32 (FFMA2 / FADD2 / FMUL2): x = s c - c, a_i + a_j and e * (a_i + a_j) on register pairs -- 13 issue
          // slots per four logits instead of 19; the MUFU / FMA-pipe exponential split is unchanged
          const uint64_t c2p = f2_pack(c2, c2), nc2p = f2_pack(-c2, -c2), aip = f2_pack(a_i, a_i);
          const bool clamp = 2.0f * c2 > 125.0f;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float4 aj = ap[i];                     // broadcast LDS.128
            const uint32_t* v = i < 8 ? v0 : v1;
            const int o = 4 * (i & 7);
            const uint64_t x01 = f2_fma(f2_pack_u(v[o], v[o + 1]), c2p, nc2p);
            const uint64_t x23 = f2_fma(f2_pack_u(v[o + 2], v[o + 3]), c2p, nc2p);
            // POLY of every 8 exponentials on the FMA pipes: 2 -> the first pair of every other group of four
            const uint64_t e01 = ((i & 1) == 0 && POLY >= 2) ? f2_ex2_fma(x01, clamp) : f2_ex2(x01);
            const uint64_t e23 = ((i & 1) == 0 && POLY >= 4) ? f2_ex2_fma(x23, clamp) : f2_ex2(x23);
            float h0, h1, h2, h3;
            f2_unpack(f2_mul(e01, f2_add(aip, f2_pack(aj.x, aj.y))), h0, h1);
            f2_unpack(f2_mul(e23, f2_add(aip, f2_pack(aj.z, aj.w))), h2, h3);
            h[2 * i] = pack_bf16x2(h0, h1);
            h[2 * i + 1] = pack_bf16x2(h2, h3);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float x[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int c = 2 * i + u;
              const int col = cb + c;
              const float sv = __uint_as_float(c < 32 ? v0[c & 31] : v1[c & 31]);
              float xv = 0.f;
              if (valid && col < p.m_cols && col != g) {
                if (col == pj) xv = gp_i + (kWait ? __ldcg(p.gpos_c + (size_t)col * p.cstride) : p.gpos_c[(size_t)col * p.cstride]);
                else xv = ex2(fmaf(sv, c2, -c2)) * (a_i + as_[c]);
              }
              x[u] = xv;
            }
            h[i] = pack_bf16x2(x[0], x[1]);
          }
        }
        // H (packed bf16 pairs) goes to columns [32 hc, 32 hc + 32) of the stage: S columns this thread has consumed
        tmem_st_x32(tbase + (uint32_t)hc * 32u, h);
      }
      if ((warp == 2 || (!COLSPLIT && warp == 6)) && lane == 0) SM3_TR(5, it);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_hfull(as));
      if ((warp == 2 || (!COLSPLIT && warp == 6)) && lane == 0) SM3_TR(6, it);
    }

    // ---- epilogue: dZ (TMEM fp32) * 1/T -> global partial ----
    mbar_wait(bar_dzfull, 0, bar_limit);
    tc_fence_after();
    if (warp == 2 && lane == 0) SM3_TR(7, 3);                              // trace: dZ complete, epilogue starts
    float* dst = p.dz_partial + ((size_t)split * p.m_rows + (valid ? l : 0)) * D;
#pragma unroll
    for (int ch = 0; ch < 2 * DP; ++ch) {
      if ((ch & 1) == grp) {
        uint32_t v[32];
        tmem_ld_x32(tmem + lane_addr + kColDZ + ch * 32, v);
        tmem_ld_wait(v);
        if (valid) {
          float4* o = reinterpret_cast<float4*>(dst + ch * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i)
            o[i] = make_float4(__uint_as_float(v[4 * i]) * p.inv_T, __uint_as_float(v[4 * i + 1]) * p.inv_T,
                               __uint_as_float(v[4 * i + 2]) * p.inv_T, __uint_as_float(v[4 * i + 3]) * p.inv_T);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) SM3_TR(7, 4);                                      // trace: CTA end
  if (warp == 1) tmem_dealloc(tmem, 512);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
// column splits: minimise  waves(row_tiles * s) * (tiles_per_split + setup)  on this many SMs
int tc_pick_splits(int row_tiles, int col_tiles, int setup_tiles, int max_splits) {
  const int sms = num_sms();
  int best = 1;
  double best_cost = 1e30;
  const int smax = col_tiles < max_splits ? col_tiles : max_splits;
  for (int s = 1; s <= smax; ++s) {
    const int tps = (col_tiles + s - 1) / s;
    const int real = (col_tiles + tps - 1) / tps;
    if (real != s) continue;
    const long ctas = (long)row_tiles * s;
    const long waves = (ctas + sms - 1) / sms;
    const double cost = (double)waves * (tps + setup_tiles) + 0.01 * s;
    if (cost < best_cost) { best_cost = cost; best = s; }
  }
  return best;
}

struct TcPlan {
  int row_tiles, col_tiles, splits, tiles_per_split, bm;
};
// forward: 256-row CTAs (half the L2 traffic) once there are enough rows to fill the machine that way
int tc_bwd_bn(int dp);      // backward column-tile width (defined with the other knobs below)
int tc_fwd_rows_per_cta(int m_rows, int m_cols) {
  const char* e = getenv("SM3_TC_FWD_BM");
  if (e) return atoi(e) == 128 ? 128 : 256;
  const long work = (long)((m_rows + 255) / 256) * ((m_cols + 127) / 128);     // (256-row block, column tile) pairs
  return work >= 8L * num_sms() ? 256 : 128;
}
TcPlan tc_plan(const InfoNceProblem& pb, bool bwd) {
  TcPlan pl;
  const int m_rows = 2 * pb.n_local, m_cols = 2 * pb.n_global;
  const int bn = bwd ? tc_bwd_bn(pb.D / 64) : 128;
  pl.bm = bwd ? kBM : tc_fwd_rows_per_cta(m_rows, m_cols);
  pl.row_tiles = (m_rows + pl.bm - 1) / pl.bm;
  pl.col_tiles = (m_cols + bn - 1) / bn;
  if (pb.skip_local) pl.col_tiles -= 2 * (pb.n_local / bn);
  int max_splits = 32;
  if (bwd) {   // bound the fp32 partial-gradient workspace to ~1 GiB
    const size_t per = (size_t)m_rows * pb.D * 4;
    const size_t cap = ((size_t)1 << 30) / (per ? per : 1);
    if ((size_t)max_splits > cap) max_splits = cap < 1 ? 1 : (int)cap;
  }
  pl.splits = tc_pick_splits(pl.row_tiles, pl.col_tiles, bwd ? (bn == 128 ? 4 : 6) : (pl.bm == 256 ? 3 : 4), max_splits);
  if (!bwd && pb.push_mode) {
    // owner-ordered tiles: every split takes every S-th tile of each owner's range, so S must divide the range length
    pl.bm = 256;
    pl.row_tiles = (m_rows + 255) / 256;
    const int L = pb.n_local / 128;
    const int sms = num_sms();
    int best = 1;
    double best_cost = 1e30;
    for (int sdiv = 1; sdiv <= L && sdiv <= max_splits; ++sdiv) {
      if (L % sdiv) continue;
      const long waves = ((long)pl.row_tiles * sdiv + sms - 1) / sms;
      const double cost = (double)waves * (pl.col_tiles / sdiv + 3) + 0.01 * sdiv;
      if (cost < best_cost) { best_cost = cost; best = sdiv; }
    }
    pl.splits = best;
    pl.tiles_per_split = pl.col_tiles / best;
    return pl;
  }
  if (const char* e = getenv(bwd ? "SM3_TC_BWD_SPLITS" : "SM3_TC_FWD_SPLITS")) {     // tuning override (sweeps)
    const int want = atoi(e);
    if (want >= 1 && want <= max_splits && want <= pl.col_tiles) {
      const int tps = (pl.col_tiles + want - 1) / want;
      pl.splits = (pl.col_tiles + tps - 1) / tps;          // no empty split
    }
  }
  pl.tiles_per_split = (pl.col_tiles + pl.splits - 1) / pl.splits;
  return pl;
}

void fill_params(const InfoNceProblem& pb, const TcPlan& pl, TcParams& p, int bn) {
  if (pb.skip_local) {
    p.skip_len = pb.n_local / bn;
    p.skip_a = pb.pair_offset / bn;
    p.skip_b = (pb.n_global + pb.pair_offset) / bn;
  } else {
    p.skip_len = 0;
    p.skip_a = p.skip_b = 0x7fffffff;
  }
  p.wait_flags = pb.wait_flags; p.wait_world = pb.wait_world; p.wait_channel = pb.wait_channel;
  p.wait_epoch = pb.wait_epoch;
  p.wait_timeout_ns = pb.wait_flags != nullptr ? peer_timeout_ns() : 0ull;
  p.loc_len = pb.n_local / bn;                      // only consulted when wait_flags != nullptr (n_local % 128 == 0)
  p.loc_a = pb.pair_offset / bn;
  p.loc_b = (pb.n_global + pb.pair_offset) / bn;
  p.n_local = pb.n_local; p.pair_offset = pb.pair_offset; p.n_global = pb.n_global; p.D = pb.D;
  p.m_rows = 2 * pb.n_local; p.m_cols = 2 * pb.n_global;
  p.inv_T = pb.inv_T; p.c2 = pb.inv_T * kLog2eTC;
  p.tiles_per_split = pl.tiles_per_split; p.col_tiles = pl.col_tiles;
  p.z_rows = (const __nv_bfloat16*)pb.z_rows;
  p.own_order = 0; p.own_L = 1; p.own_S = 1; p.push_src = nullptr; p.push_vpr = 0; p.push_ctas = 0; p.push_counter = nullptr;
  p.push_dst.world = 0; p.push_flags.world = 0;
}

// number of softmax warp groups (8 warps each): tuning knob, SM3_TC_GROUPS=1|2.  Measured on B200 (cfg4): one group
// is faster in the forward (1.93 vs 2.57 ms: with 3 TMEM S stages two tiles in flight starve the MMA warp of a free
// stage) and equal in the backward, so 1 is the default.
// The tuning knobs are read from the environment once and cached; sm3_debug_reload_env() drops the cache so that a
// sweep (tools/tc_sweep.py) can change them inside one process.
int g_knob_groups = 0, g_knob_poly = -1, g_knob_bwd_ns = -1, g_knob_bwd_v = -1;
int tc_groups() {
  int& g = g_knob_groups;
  if (g == 0) {
    const char* e = getenv("SM3_TC_GROUPS");
    g = (e && e[0] == '2') ? 2 : 1;
  }
  return g;
}

// exponentials per 8 evaluated on the FMA pipes (ex2_fma) instead of MUFU in the forward: tuning knob
// SM3_TC_POLY=0..4.  Default 0: on B200 the kernel runs under the 1 kW power cap (SM clock ~1.3-1.5 GHz under this
// load), where trading one MUFU op for ~11 FMA/ALU ops shortens the cycle count per tile (trace build: 1550 -> 1350)
// but not the wall time; measured 1.74 / 1.78 / 1.80 ms for POLY = 0 / 2 / 3 at cfg4.
int tc_poly(int dp) {
  int& v = g_knob_poly;
  if (v == -1) {
    const char* e = getenv("SM3_TC_POLY");
    v = e ? atoi(e) : -2;                       // -2: not set -> by embedding width
    if (v != -2 && (v < 0 || v > 4)) v = 0;
  }
  // D <= 128: the forward is MUFU-bound (1024 MUFU cycles vs 512 MMA cycles per 128 x 128 tile), so a quarter of the
  // exponentials go to the FMA pipes by default; D > 128: MMA and MUFU are level and the chip is power-capped -> 0.
  if (v == -2) return dp <= 2 ? 2 : 0;
  return v;
}

template <int DP, int NG, int POLY>
int launch_fwd_ng(const CUtensorMap& tmap, const TcParams& p, const TcPlan& pl, cudaStream_t st) {
  SM3_SMEM_ATTR_ONCE((infonce_tc_fwd_kernel<DP, NG, POLY>), FwdCfg<DP>::SMEM);
  SM3_CHECK_CUDA(launch_k(infonce_tc_fwd_kernel<DP, NG, POLY>, dim3(pl.row_tiles, pl.splits), dim3(64 + 256 * NG), FwdCfg<DP>::SMEM, st, tmap, p));
  SM3_CHECK_CUDA(cudaGetLastError());
  return SM3_OK;
}
template <int DP, int POLY>
int launch_fwd2(const CUtensorMap& tmap, const TcParams& p, const TcPlan& pl, cudaStream_t st) {
  if (p.push_src != nullptr) {
    SM3_SMEM_ATTR_ONCE((infonce_tc_fwd2_kernel<DP, POLY, true>), FwdCfg<DP>::SMEM);
    SM3_CHECK_CUDA(launch_k(infonce_tc_fwd2_kernel<DP, POLY, true>, dim3(pl.row_tiles, pl.splits), dim3(384), FwdCfg<DP>::SMEM, st, tmap, p));
    return SM3_OK;
  }
  SM3_SMEM_ATTR_ONCE((infonce_tc_fwd2_kernel<DP, POLY, false>), FwdCfg<DP>::SMEM);
  SM3_CHECK_CUDA(launch_k(infonce_tc_fwd2_kernel<DP, POLY, false>, dim3(pl.row_tiles, pl.splits), dim3(320), FwdCfg<DP>::SMEM, st, tmap, p));
  SM3_CHECK_CUDA(cudaGetLastError());
  return SM3_OK;
}
template <int DP>
int launch_fwd(const CUtensorMap& tmap, const TcParams& p, const TcPlan& pl, cudaStream_t st) {
  if (pl.bm == 256) return tc_poly(DP) == 2 ? launch_fwd2<DP, 2>(tmap, p, pl, st) : launch_fwd2<DP, 0>(tmap, p, pl, st);
  if (tc_groups() == 2) {
    switch (tc_poly(DP)) {
      case 0: return launch_fwd_ng<DP, 2, 0>(tmap, p, pl, st);
      case 3: return launch_fwd_ng<DP, 2, 3>(tmap, p, pl, st);
      default: return launch_fwd_ng<DP, 2, 2>(tmap, p, pl, st);
    }
  }
  switch (tc_poly(DP)) {
    case 0: return launch_fwd_ng<DP, 1, 0>(tmap, p, pl, st);
    case 1: return launch_fwd_ng<DP, 1, 1>(tmap, p, pl, st);
    case 3: return launch_fwd_ng<DP, 1, 3>(tmap, p, pl, st);
    case 4: return launch_fwd_ng<DP, 1, 4>(tmap, p, pl, st);
    default: return launch_fwd_ng<DP, 1, 2>(tmap, p, pl, st);
  }
}
template <int DP, int NG, bool kWait, int NS>
int launch_bwd_ng(const CUtensorMap& tmap, const TcParams& p, const TcPlan& pl, cudaStream_t st) {
  SM3_SMEM_ATTR_ONCE((infonce_tc_bwd_kernel<DP, NG, kWait, NS>), (BwdCfg<DP, NS>::SMEM));
  SM3_CHECK_CUDA(launch_k(infonce_tc_bwd_kernel<DP, NG, kWait, NS>, dim3(pl.row_tiles, pl.splits), dim3(64 + 256 * NG), BwdCfg<DP, NS>::SMEM, st, tmap, p));
  SM3_CHECK_CUDA(cudaGetLastError());
  return SM3_OK;
}
// S/H stages of the backward: 4 for D <= 128 (latency-bound chain, TMEM has room), 2 otherwise; SM3_TC_BWD_NS=2 forces 2.
int tc_bwd_stages(int dp) {
  int& forced = g_knob_bwd_ns;
  if (forced < 0) {
    const char* e = getenv("SM3_TC_BWD_NS");
    forced = (e && e[0] == '2') ? 2 : 0;
  }
  if (forced == 2 || dp > 2) return 2;
  return 4;
}
template <int DP, bool kWait, int NS, int BNT, int POLY, bool COLSPLIT = false>
int launch_bwd2(const CUtensorMap& tmap, const TcParams& p, const TcPlan& pl, cudaStream_t st) {
  using C = Bwd2Cfg<DP, NS, BNT>;
  SM3_SMEM_ATTR_ONCE((infonce_tc_bwd2_kernel<DP, kWait, NS, BNT, POLY, COLSPLIT>), (C::SMEM));
  SM3_CHECK_CUDA(launch_k(infonce_tc_bwd2_kernel<DP, kWait, NS, BNT, POLY, COLSPLIT>, dim3(pl.row_tiles, pl.splits), dim3(320), C::SMEM, st, tmap, p));
  SM3_CHECK_CUDA(cudaGetLastError());
  return SM3_OK;
}
// Which backward kernel (SM3_TC_BWD_V overrides; default by embedding width):
//   1 = first form (softmax warps split each 64-column tile)              -- default for D > 128
//   2 = tile-alternating groups + a_j through shared memory, 64-column tiles
//   3 = the same with 128-COLUMN tiles (D <= 128 only: TMEM has room for two 128-column S/H stages)
//   4 = 128-column tiles, all eight softmax warps on every tile, the two groups split its columns -- default for D <= 128.
// Measured with the trace build on B200 (cfg2, D = 128): the thread that issues tcgen05.mma needs ~52 cycles per
// instruction whatever its size, so the N = 64 S-MMAs of a 64-column tile (32 tensor cycles each) cannot keep the pipe
// busy: 12 issues = ~940 cycles per tile against 512 cycles of tensor work.  128-column tiles halve the issues per column.
int tc_bwd_version(int dp) {
  int& v = g_knob_bwd_v;
  if (v < 0) {
    const char* e = getenv("SM3_TC_BWD_V");
    v = (e && e[0] >= '1' && e[0] <= '4') ? e[0] - '0' : 0;
  }
  int r = v != 0 ? v : (dp <= 2 ? 4 : 1);
  if (r >= 3 && dp > 2) r = 2;
  return r;
}
int tc_bwd_bn(int dp) { return tc_bwd_version(dp) >= 3 ? 128 : 64; }
// exponentials per 8 on the FMA pipes in the 128-column backward (SM3_TC_BWD_POLY=0|2; default 2: at D <= 128 the
// softmax side is MUFU-bound, 1024 MUFU cycles per 128 x 128 tile against 1024 tensor cycles)
int g_knob_bwd_poly = -1;
int tc_bwd_poly() {
  int& v = g_knob_bwd_poly;
  if (v < 0) {
    const char* e = getenv("SM3_TC_BWD_POLY");
    v = e ? (atoi(e) == 2 ? 2 : 0) : 2;
  }
  return v;
}
template <int DP>
int launch_bwd(const CUtensorMap& tmap, const TcParams& p, const TcPlan& pl, cudaStream_t st) {
  const int ver = tc_bwd_version(DP);
  if constexpr (DP <= 2) {
    if (ver == 3) {
      const bool poly = tc_bwd_poly() == 2;
      if (p.wait_flags != nullptr)
        return poly ? launch_bwd2<DP, true, 2, 128, 2>(tmap, p, pl, st) : launch_bwd2<DP, true, 2, 128, 0>(tmap, p, pl, st);
      return poly ? launch_bwd2<DP, false, 2, 128, 2>(tmap, p, pl, st) : launch_bwd2<DP, false, 2, 128, 0>(tmap, p, pl, st);
    }
    if (ver == 4) {
      const bool poly = tc_bwd_poly() == 2;
      if (p.wait_flags != nullptr)
        return poly ? launch_bwd2<DP, true, 2, 128, 2, true>(tmap, p, pl, st) : launch_bwd2<DP, true, 2, 128, 0, true>(tmap, p, pl, st);
      return poly ? launch_bwd2<DP, false, 2, 128, 2, true>(tmap, p, pl, st) : launch_bwd2<DP, false, 2, 128, 0, true>(tmap, p, pl, st);
    }
  }
  if (ver >= 2) {
    constexpr int NS2 = DP <= 2 ? 4 : 2;
    const bool two = tc_bwd_stages(DP) == 2;
    if (p.wait_flags != nullptr)
      return two ? launch_bwd2<DP, true, 2, 64, 0>(tmap, p, pl, st) : launch_bwd2<DP, true, NS2, 64, 0>(tmap, p, pl, st);
    return two ? launch_bwd2<DP, false, 2, 64, 0>(tmap, p, pl, st) : launch_bwd2<DP, false, NS2, 64, 0>(tmap, p, pl, st);
  }
  // (the in-kernel-wait form of the multi-rank fused exchange stays on 2 stages until it has been run on >= 2 GPUs)
  if (p.wait_flags != nullptr) return launch_bwd_ng<DP, 1, true, 2>(tmap, p, pl, st);
  if constexpr (DP <= 2) {
    if (tc_bwd_stages(DP) == 4)
      return tc_groups() == 1 ? launch_bwd_ng<DP, 1, false, 4>(tmap, p, pl, st) : launch_bwd_ng<DP, 2, false, 4>(tmap, p, pl, st);
  }
  return tc_groups() == 1 ? launch_bwd_ng<DP, 1, false, 2>(tmap, p, pl, st) : launch_bwd_ng<DP, 2, false, 2>(tmap, p, pl, st);
}

size_t bwd_acol_offset(const InfoNceProblem& pb, const TcPlan& pl) {
  const size_t dz = (size_t)pl.splits * 2 * pb.n_local * pb.D * sizeof(float);
  return (dz + 255) & ~(size_t)255;
}


// symmetric forward: when it applies and how the flat tile list is cut (see kSymFlag in common.cuh).  SM3_TC_FWD_SYM=0
// keeps the full-matrix kernel, 2 uses the symmetric one at every size it supports.
int g_knob_sym = -1;
struct SymPlan {
  bool on;
  int T, P, tpc, nctas, maxseg;
  long W;
  size_t ws;
};
SymPlan tc_sym_plan(const InfoNceProblem& pb) {
  SymPlan sp{};
  if (g_knob_sym < 0) {
    const char* e = getenv("SM3_TC_FWD_SYM");
    g_knob_sym = (e && e[0] == '0') ? 0 : (e && e[0] == '2') ? 2 : 1;      // 2: also below the size threshold (tests)
  }
  if (!g_knob_sym || pb.n_local != pb.n_global || pb.pair_offset != 0 || pb.skip_local || pb.wait_flags != nullptr ||
      pb.push_mode || pb.extra_neg_sum != nullptr || pb.z_rows != pb.z_cols)
    return sp;
  const long M = 2L * pb.n_global;
  if (M % 256 != 0) return sp;
  sp.T = (int)(M / 128);
  sp.P = sp.T / 2;
  sp.W = (long)sp.P * (sp.P + 1);
  const int sms = num_sms();
  if (g_knob_sym != 2 && sp.W < 2L * sms) return sp;   // too few tiles to fill the machine with half of them
  sp.tpc = (int)((sp.W + sms - 1) / sms);
  sp.nctas = (int)((sp.W + sp.tpc - 1) / sp.tpc);
  sp.maxseg = sym_maxseg(sp.T, sp.tpc);
  sp.ws = (size_t)(sp.maxseg + sp.P) * (size_t)M * sizeof(float) + 256;
  // in the fused step the loss kernel reads these sums while it writes a_j behind the backward's partial-gradient slabs
  if (sp.ws > bwd_acol_offset(pb, tc_plan(pb, true))) return sp;
  sp.on = true;
  return sp;
}
int g_knob_sym_nq = -1;
template <int DP, int POLY, int NQ>
int launch_fwdsym_q(const CUtensorMap& tmap, const TcParams& p, const SymPlan& sp, cudaStream_t st) {
  SM3_SMEM_ATTR_ONCE((infonce_tc_fwdsym_kernel<DP, POLY, NQ>), SymCfg<DP>::SMEM);
  SM3_CHECK_CUDA(launch_k(infonce_tc_fwdsym_kernel<DP, POLY, NQ>, dim3(sp.nctas), dim3(64 + 128 * NQ), SymCfg<DP>::SMEM, st, tmap, p));
  SM3_CHECK_CUDA(cudaGetLastError());
  return SM3_OK;
}
// softmax warps per lane quadrant: SM3_TC_SYM_NQ = 2 | 4.  Sixteen softmax warps (4) do not beat eight (2): 1.05-1.23 ms
// against 1.00-1.05 ms at cfg4, 32.7-36.1 against 34.6-35.7 us at cfg2 (the 16x256b TMEM loads queue behind each other
// and the per-warp fixed work -- barrier waits, butterfly, fold -- doubles), so 2 is the default.
template <int DP, int POLY>
int launch_fwdsym_p(const CUtensorMap& tmap, const TcParams& p, const SymPlan& sp, cudaStream_t st) {
  if (g_knob_sym_nq < 0) {
    const char* e = getenv("SM3_TC_SYM_NQ");
    g_knob_sym_nq = (e && e[0] == '4') ? 4 : 2;
  }
  return g_knob_sym_nq == 2 ? launch_fwdsym_q<DP, POLY, 2>(tmap, p, sp, st) : launch_fwdsym_q<DP, POLY, 4>(tmap, p, sp, st);
}
// SM3_TC_POLY = 0 | 2 | 4 of every 8 exponentials of the symmetric kernel leave the MUFU pipe for the FMA pipes (packed
// FFMA2 / FADD2 form, ~5.5 issue slots per exponential).  Measured on B200 (profiles/r02_sym_variants.txt): cfg4
// 1.05 / 1.00 / 1.13 ms and cfg2 34.6 / 35.7 / 35.6 us for 0 / 2 / 4 -- within run-to-run noise of each other, because
// the kernel is bound by neither pipe (MUFU 24-50 % busy, issue slots ~40 %) but by the latency of its per-tile chain
// S ready -> tcgen05.ld -> exp -> column butterfly with only two softmax warps per sub-partition.  Default 2.
template <int DP>
int launch_fwdsym(const CUtensorMap& tmap, const TcParams& p, const SymPlan& sp, cudaStream_t st) {
  (void)tc_poly(DP);                        // reads SM3_TC_POLY into g_knob_poly (-2 = not set)
  const int poly = g_knob_poly == -2 ? 2 : g_knob_poly;
  if (poly >= 3) return launch_fwdsym_p<DP, 4>(tmap, p, sp, st);
  if (poly >= 1) return launch_fwdsym_p<DP, 2>(tmap, p, sp, st);
  return launch_fwdsym_p<DP, 0>(tmap, p, sp, st);
}

}  // namespace

size_t infonce_tc_acol_offset(const InfoNceProblem& pb) { return bwd_acol_offset(pb, tc_plan(pb, true)); }

bool infonce_tc_push_supported(const InfoNceProblem& pb) {
  return infonce_tc_supported(pb) && pb.n_local % 128 == 0 && pb.n_global > pb.n_local &&
         tc_fwd_rows_per_cta(2 * pb.n_local, 2 * pb.n_global) == 256;
}

bool infonce_tc_supported(const InfoNceProblem& pb) {
  static int sm100 = -1;
  if (sm100 < 0) sm100 = sm3_device_supported() == 1 ? 1 : 0;
  return sm100 == 1 && pb.dtype == SM3_BF16 && pb.D % 64 == 0 && pb.D >= 64 && pb.D <= 256 && aligned16(pb.z_rows) &&
         aligned16(pb.z_cols);
}

size_t infonce_tc_workspace(const InfoNceProblem& pb, int backward) {
  if (pb.D % 64 != 0 || pb.D < 64 || pb.D > 256) return 0;
  const TcPlan pl = tc_plan(pb, backward != 0);
  if (!backward) {
    const size_t full = (size_t)pl.splits * 2 * pb.n_local * sizeof(float) + 256;
    const SymPlan sp = tc_sym_plan(pb);
    return sp.on && sp.ws > full ? sp.ws : full;
  }
  const size_t m_pad = ((size_t)2 * pb.n_global + 127) / 128 * 128 + 128;   // whole 128-column tiles + one of slack
  return bwd_acol_offset(pb, pl) + m_pad * sizeof(float) + 256;
}

int infonce_tc_fwd(const InfoNceProblem& pb, float* pos, float* lse_neg, float* neg_sum, void* ws, size_t ws_bytes,
                   cudaStream_t st) {
  SM3_REQUIRE(pb.n_local >= 1 && pb.n_global >= pb.n_local && pb.pair_offset >= 0 &&
                  pb.pair_offset + pb.n_local <= pb.n_global && pb.n_global <= (1 << 29),
              SM3_ERR_SHAPE, "infonce: bad row block (n_local=%d offset=%d n_global=%d)", pb.n_local, pb.pair_offset,
              pb.n_global);
  SM3_REQUIRE(!pb.skip_local || (pb.n_local % 128 == 0 && pb.n_global > pb.n_local), SM3_ERR_SHAPE,
              "infonce(tc): skip_local needs n_local %% 128 == 0 and more than one rank");
  SM3_REQUIRE(pb.wait_flags == nullptr || (pb.n_local % 128 == 0 && !pb.skip_local), SM3_ERR_SHAPE,
              "infonce(tc): in-kernel peer waits need n_local %% 128 == 0");
  SM3_REQUIRE(ws_bytes >= infonce_tc_workspace(pb, 0), SM3_ERR_WORKSPACE, "infonce(tc) fwd: workspace too small");
  const TcPlan pl = tc_plan(pb, false);
  TcParams p{};
  fill_params(pb, pl, p, 128);
  p.pos = pos;
  p.partial = (float*)ws;
  if (pb.push_mode) {
    SM3_REQUIRE(pb.wait_flags != nullptr && pb.n_local % 128 == 0 && pl.bm == 256 && !pb.skip_local, SM3_ERR_SHAPE,
                "infonce(tc): push mode needs the fused exchange, n_local %% 128 == 0 and the 256-row kernel");
    p.own_order = 1; p.own_L = pb.n_local / 128; p.own_S = pl.splits;
    if (pb.push_src != nullptr) {
      SM3_REQUIRE(pb.push != nullptr && pb.push->counter != nullptr && pb.push->data.world == pb.wait_world &&
                      pb.push->flags.world == pb.wait_world, SM3_ERR_SHAPE, "infonce(tc): push mode without destinations");
      p.push_src = (const uint4*)pb.push_src;
      p.push_vpr = pb.D * 2 / 16;
      const int ctas = pl.row_tiles * pl.splits;
      p.push_ctas = ctas < num_sms() ? ctas : num_sms();
      p.push_dst = pb.push->data; p.push_flags = pb.push->flags; p.push_counter = pb.push->counter;
    }
  }
  CUtensorMap tmap;
  int rc = make_tmap_bf16(&tmap, pb.z_cols, (uint64_t)p.m_cols, (uint64_t)pb.D, 128);
  if (rc) return rc;
  const SymPlan sp = tc_sym_plan(pb);
  if (sp.on) {
    p.sym_T = sp.T; p.sym_tpc = sp.tpc; p.sym_maxseg = sp.maxseg; p.sym_W = sp.W;
    switch (pb.D / 64) {
      case 1: rc = launch_fwdsym<1>(tmap, p, sp, st); break;
      case 2: rc = launch_fwdsym<2>(tmap, p, sp, st); break;
      case 3: rc = launch_fwdsym<3>(tmap, p, sp, st); break;
      default: rc = launch_fwdsym<4>(tmap, p, sp, st); break;
    }
    if (rc) return rc;
    const int enc = kSymFlag | sp.tpc;
    if (pb.no_finalize) return enc;
    return infonce_finalize_launch(p.partial, enc, p.m_rows, pb.inv_T, neg_sum, lse_neg, st, nullptr);
  }
  switch (pb.D / 64) {
    case 1: rc = launch_fwd<1>(tmap, p, pl, st); break;
    case 2: rc = launch_fwd<2>(tmap, p, pl, st); break;
    case 3: rc = launch_fwd<3>(tmap, p, pl, st); break;
    default: rc = launch_fwd<4>(tmap, p, pl, st); break;
  }
  if (rc) return rc;
  if (pb.no_finalize) return pl.splits;             // the caller folds the [splits][m_rows] partial sums itself
  return infonce_finalize_launch(p.partial, pl.splits, p.m_rows, pb.inv_T, neg_sum, lse_neg, st, pb.extra_neg_sum);
}

// ---- cross-rank symmetric forward: plan, workspace, launch (see MrPlan in common.cuh) ----
MrPlan infonce_tc_mr_plan(int n_local, int world, int rank) {
  MrPlan m{};
  if (world < 2 || world > 16 || n_local % 128 != 0 || n_local < 128) return m;
  m.W = world; m.rank = rank;
  m.H = (world - 1) / 2;
  m.anti = (world % 2 == 0) ? (rank < world / 2 ? 1 : 2) : 0;
  m.np = m.H + (m.anti ? 1 : 0);
  m.T_l = 2 * n_local / 128;
  m.P_l = m.T_l / 2;
  m.flat = mr_prefix(m, m.P_l);
  const int sms = num_sms();
  m.tpc = (int)((m.flat + sms - 1) / sms);
  if (m.tpc < 1) m.tpc = 1;
  m.nctas = (int)((m.flat + m.tpc - 1) / m.tpc);
  int cmax = 0;
  for (int R = 0; R < m.P_l; ++R) cmax = mr_count(m, R) > cmax ? mr_count(m, R) : cmax;
  m.maxseg = (cmax - 1) / m.tpc + 2;
  m.on = 1;
  return m;
}
size_t infonce_tc_mr_workspace(const MrPlan& m) {
  return (size_t)(m.maxseg + (1 + m.np) * m.P_l) * (size_t)m.T_l * 128 * sizeof(float) + 256;
}
size_t infonce_tc_mr_workspace_any(int n_local, int world) {
  size_t w = 0;
  for (int r = 0; r < world; ++r) {
    const MrPlan m = infonce_tc_mr_plan(n_local, world, r);
    if (!m.on) return 0;
    const size_t b = infonce_tc_mr_workspace(m);
    w = b > w ? b : w;
  }
  return w;
}
namespace {
template <int DP>
int launch_fwdsym_mr(const CUtensorMap& tmap, const TcParams& p, cudaStream_t st) {
  (void)tc_poly(DP);
  const int poly = g_knob_poly == -2 ? 2 : g_knob_poly;
  if (poly >= 1) {
    SM3_SMEM_ATTR_ONCE((infonce_tc_fwdsym_mr_kernel<DP, 2, 2>), SymCfg<DP>::SMEM);
    SM3_CHECK_CUDA(launch_k(infonce_tc_fwdsym_mr_kernel<DP, 2, 2>, dim3(p.mr.nctas), dim3(320), SymCfg<DP>::SMEM, st, tmap, p));
  } else {
    SM3_SMEM_ATTR_ONCE((infonce_tc_fwdsym_mr_kernel<DP, 0, 2>), SymCfg<DP>::SMEM);
    SM3_CHECK_CUDA(launch_k(infonce_tc_fwdsym_mr_kernel<DP, 0, 2>, dim3(p.mr.nctas), dim3(320), SymCfg<DP>::SMEM, st, tmap, p));
  }
  SM3_CHECK_CUDA(cudaGetLastError());
  return SM3_OK;
}
}  // namespace
// pb: this rank's row block against the all-gathered columns (wait_flags etc. as for the fused exchange).  Leaves the row
// sums / own-block column sums / partner column slabs in ws (infonce_tc_mr_workspace bytes) and the positives in pos.
int infonce_tc_fwd_mr(const InfoNceProblem& pb, const MrPlan& m, float* pos, void* ws, size_t ws_bytes, cudaStream_t st) {
  SM3_REQUIRE(m.on && pb.wait_flags != nullptr && pb.n_local % 128 == 0 && pb.n_global == pb.n_local * m.W &&
                  pb.pair_offset == m.rank * pb.n_local, SM3_ERR_SHAPE, "infonce(tc) symmetric exchange: bad plan");
  SM3_REQUIRE(ws_bytes >= infonce_tc_mr_workspace(m), SM3_ERR_WORKSPACE, "infonce(tc) symmetric exchange: workspace too small");
  const TcPlan pl = tc_plan(pb, false);
  TcParams p{};
  fill_params(pb, pl, p, 128);
  p.pos = pos;
  p.partial = (float*)ws;
  p.mr = m;
  CUtensorMap tmap;
  int rc = make_tmap_bf16(&tmap, pb.z_cols, (uint64_t)p.m_cols, (uint64_t)pb.D, 128);
  if (rc) return rc;
  switch (pb.D / 64) {
    case 1: return launch_fwdsym_mr<1>(tmap, p, st);
    case 2: return launch_fwdsym_mr<2>(tmap, p, st);
    case 3: return launch_fwdsym_mr<3>(tmap, p, st);
    default: return launch_fwdsym_mr<4>(tmap, p, st);
  }
}

int infonce_tc_bwd(const InfoNceProblem& pb, const float* gpos_r, const float* glse_r, const float* nsum_r,
                   const float* gpos_c, const float* glse_c, const float* nsum_c, void* ws, size_t ws_bytes,
                   cudaStream_t st) {
  SM3_REQUIRE(pb.n_local >= 1 && pb.n_global >= pb.n_local && pb.pair_offset >= 0 &&
                  pb.pair_offset + pb.n_local <= pb.n_global && pb.n_global <= (1 << 29),
              SM3_ERR_SHAPE, "infonce: bad row block (n_local=%d offset=%d n_global=%d)", pb.n_local, pb.pair_offset,
              pb.n_global);
  SM3_REQUIRE(!pb.skip_local || (pb.n_local % 128 == 0 && pb.n_global > pb.n_local), SM3_ERR_SHAPE,
              "infonce(tc): skip_local needs n_local %% 128 == 0 and more than one rank");
  SM3_REQUIRE(ws_bytes >= infonce_tc_workspace(pb, 1), SM3_ERR_WORKSPACE, "infonce(tc) bwd: workspace too small");
  const TcPlan pl = tc_plan(pb, true);
  TcParams p{};
  const int bn = tc_bwd_bn(pb.D / 64);
  fill_params(pb, pl, p, bn);
  p.gpos_r = gpos_r; p.glse_r = glse_r; p.nsum_r = nsum_r; p.gpos_c = gpos_c; p.cstride = pb.col_stride;
  p.dz_partial = (float*)ws;
  SM3_REQUIRE(pb.wait_flags == nullptr || (pb.n_local % 128 == 0 && !pb.skip_local && pb.acol_direct != nullptr),
              SM3_ERR_SHAPE, "infonce(tc): in-kernel peer waits need n_local %% 128 == 0 and materialised column statistics");
  if (pb.acol_direct != nullptr) {
    p.acol = pb.acol_direct;                        // written by the owner ranks (fused exchange): no prep kernel
  } else {
    float* acol = (float*)((char*)ws + bwd_acol_offset(pb, pl));
    p.acol = acol;
    const int m_pad = (p.m_cols + 127) / 128 * 128 + 128;
    tc_bwd_prep_kernel<<<(m_pad + 255) / 256, 256, 0, st>>>(glse_c, nsum_c, pb.col_stride, p.m_cols, m_pad, acol);
    SM3_CHECK_CUDA(cudaGetLastError());
  }
  CUtensorMap tmap;
  int rc = make_tmap_bf16(&tmap, pb.z_cols, (uint64_t)p.m_cols, (uint64_t)pb.D, (uint32_t)bn);
  if (rc) return rc;
  switch (pb.D / 64) {
    case 1: rc = launch_bwd<1>(tmap, p, pl, st); break;
    case 2: rc = launch_bwd<2>(tmap, p, pl, st); break;
    case 3: rc = launch_bwd<3>(tmap, p, pl, st); break;
    default: rc = launch_bwd<4>(tmap, p, pl, st); break;
  }
  if (rc) return rc;
  return pl.splits;
}

}  // namespace sm3

// 1 when the single-rank tcgen05 forward of this shape runs the symmetric kernel; *executed_tiles (optional) receives the
// number of 128 x 128 S tiles it computes (the full-matrix kernel computes (2 n_pairs / 128)^2)
extern "C" int sm3_infonce_fwd_symmetric(int n_pairs, int D, long long* executed_tiles) {
  using namespace sm3;
  if (n_pairs < 1 || D % 64 != 0 || D < 64 || D > 256) return 0;
  InfoNceProblem pb{nullptr, nullptr, n_pairs, 0, n_pairs, D, SM3_BF16, 1.0f};
  const SymPlan sp = tc_sym_plan(pb);
  if (executed_tiles) *executed_tiles = sp.on ? 2 * sp.W : (long long)((2L * n_pairs + 127) / 128) * ((2L * n_pairs + 127) / 128);
  return sp.on ? 1 : 0;
}

// Host-side enumeration of the work lists the symmetric kernels walk, with the kernels' own index functions (no GPU
// needed): for every CTA piece of rank `rank`'s flat list, one record (cta, row pair R, global column tile, column sums
// taken 0/1, slab row the column sums go to) per visited tile.  world == 1 enumerates the single-rank symmetric kernel
// (n_local = all pairs).  Returns the number of records (<= cap written), or a negative error code.
extern "C" long long sm3_debug_sym_enumerate(int n_local, int world, int rank, int tpc_override, int* records, long long cap) {
  using namespace sm3;
  SM3_REQUIRE(n_local >= 128 && n_local % 128 == 0 && world >= 1 && world <= 16 && rank >= 0 && rank < world && records,
              SM3_ERR_SHAPE, "sym_enumerate: bad arguments");
  long long n = 0;
  auto put = [&](int cta, int R, int gtile, int docol, int slab) {
    if (n < cap) { int* r = records + 5 * n; r[0] = cta; r[1] = R; r[2] = gtile; r[3] = docol; r[4] = slab; }
    ++n;
  };
  if (world == 1) {
    const int T = 2 * n_local / 128, P = T / 2;
    const long W = (long)P * (P + 1);
    const int sms = num_sms();
    const int tpc = tpc_override > 0 ? tpc_override : (int)((W + sms - 1) / sms);
    const int nctas = (int)((W + tpc - 1) / tpc);
    for (int c = 0; c < nctas; ++c) {
      const long f0 = (long)c * tpc, f1 = (f0 + tpc < W) ? f0 + tpc : W;
      for (long f = f0; f < f1;) {
        int R, off;
        sym_decode(f, T, P, R, off);
        const long left = f1 - f;
        const int cnt = (int)((long)(T - 2 * R - off) < left ? (long)(T - 2 * R - off) : left);
        const int kseg = c - (int)(sym_flat_start(R, T, P) / tpc);
        SM3_REQUIRE(cnt > 0 && kseg >= 0 && kseg < sym_maxseg(T, tpc), SM3_ERR_SHAPE, "sym_enumerate: bad segment");
        for (int t = 0; t < cnt; ++t) put(c, R, 2 * R + off + t, (2 * R + off + t) > 2 * R + 1, R);
        f += cnt;
      }
    }
    return n;
  }
  MrPlan m = infonce_tc_mr_plan(n_local, world, rank);
  SM3_REQUIRE(m.on, SM3_ERR_SHAPE, "sym_enumerate: no plan");
  if (tpc_override > 0) {
    m.tpc = tpc_override;
    m.nctas = (int)((m.flat + m.tpc - 1) / m.tpc);
  }
  for (int c = 0; c < m.nctas; ++c) {
    const long f0 = (long)c * m.tpc, f1 = (f0 + m.tpc < m.flat) ? f0 + m.tpc : m.flat;
    for (long f = f0; f < f1;) {
      int R, off;
      mr_decode(m, f, R, off);
      const long left = f1 - f;
      const int cnt = (int)((long)(mr_count(m, R) - off) < left ? (long)(mr_count(m, R) - off) : left);
      SM3_REQUIRE(cnt > 0 && R >= 0 && R < m.P_l && off >= 0, SM3_ERR_SHAPE, "sym_enumerate: bad segment");
      for (int t = 0; t < cnt; ++t) {
        int ps, j_l;
        mr_tile(m, R, off + t, ps, j_l);
        const int owner = ps < 0 ? m.rank : mr_partner_rank(m, ps);
        put(c, R, mr_global_tile(m, owner, j_l), (ps >= 0 || j_l > 2 * R + 1) ? 1 : 0, (ps < 0 ? 0 : (1 + ps) * m.P_l) + R);
      }
      f += cnt;
    }
  }
  return n;
}

extern "C" void sm3_debug_reload_env(void) {
  sm3::g_knob_groups = 0;
  sm3::g_knob_poly = -1;
  sm3::g_knob_bwd_ns = -1;
  sm3::g_knob_bwd_v = -1;
  sm3::g_knob_bwd_poly = -1;
  sm3::g_knob_sym = -1;
  sm3::g_knob_sym_nq = -1;
}

#ifdef SM3_TRACE
extern "C" int sm3_debug_read_trace(long long* host, int n) {
  using namespace sm3;
  const int cap = kTraceTiles * kTraceKinds;
  SM3_CHECK_CUDA(cudaMemcpyFromSymbol(host, g_trace, sizeof(long long) * (n < cap ? n : cap)));
  return SM3_OK;
}
#endif

// ---------------------------------------------------------------------------------------------------
// debug probe (C ABI).  variant bit0: A operand from TMEM (else smem/TMA); bit1: B given as [k, n]
// row-major and consumed MN-major (C = A*B), else B given as [n, k] row-major, K-major (C = A*B^T).
// ---------------------------------------------------------------------------------------------------
extern "C" int sm3_debug_umma_probe(const void* a_bf16, const void* b_bf16, float* c, int n, int k, int variant,
                                    void* stream) {
  using namespace sm3;
  cudaStream_t st = (cudaStream_t)stream;
  SM3_REQUIRE(a_bf16 && b_bf16 && c, SM3_ERR_SHAPE, "probe: null pointer");
  SM3_REQUIRE(n >= 32 && n <= 256 && n % 32 == 0 && k >= 64 && k <= 256 && k % 64 == 0, SM3_ERR_SHAPE,
              "probe: need n in [32,256] %%32, k in [64,256] %%64");
  const bool a_tmem = variant & 1, b_mn = variant & 2;
  SM3_REQUIRE(!b_mn || n % 64 == 0, SM3_ERR_SHAPE, "probe: MN-major B needs n %% 64 == 0");
  CUtensorMap ta, tb;
  int rc = make_tmap_bf16(&ta, a_bf16, 128, (uint64_t)k, 128);
  if (rc) return rc;
  if (!b_mn) rc = make_tmap_bf16(&tb, b_bf16, (uint64_t)n, (uint64_t)k, (uint32_t)n);
  else rc = make_tmap_bf16(&tb, b_bf16, (uint64_t)k, (uint64_t)n, (uint32_t)k);
  if (rc) return rc;
  const int smem = 4 * 16384 + 4 * 32768 + 1024 + 64;
  const __nv_bfloat16* ag = (const __nv_bfloat16*)a_bf16;
#define SM3_PROBE(AT, BM)                                                                                           \
  do {                                                                                                              \
    SM3_CHECK_CUDA(cudaFuncSetAttribute(umma_probe_kernel<AT, BM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)); \
    umma_probe_kernel<AT, BM><<<1, 128, smem, st>>>(ta, tb, ag, c, n, k);                                           \
  } while (0)
  if (a_tmem && b_mn) SM3_PROBE(true, true);
  else if (a_tmem) SM3_PROBE(true, false);
  else if (b_mn) SM3_PROBE(false, true);
  else SM3_PROBE(false, false);
#undef SM3_PROBE
  SM3_CHECK_CUDA(cudaGetLastError());
  return SM3_OK;
}

// debug microbenchmark (C ABI): see umma_rate_kernel.  out_host[0] = issue cycles, [1] = total cycles, [3..] = tcgen05.ld
// round trips completed by the contending warps.
extern "C" int sm3_debug_umma_rate(int n, int a_from_tmem, int count, int ldtm_warps, long long* out_host) {
  using namespace sm3;
  SM3_REQUIRE(n >= 16 && n <= 256 && n % 16 == 0 && count >= 1 && count <= 4096 && ldtm_warps >= 0 && ldtm_warps <= 8 && out_host,
              SM3_ERR_SHAPE, "umma_rate: bad arguments");
  long long* dev = nullptr;
  SM3_CHECK_CUDA(cudaMalloc(&dev, 16 * sizeof(long long)));
  SM3_CHECK_CUDA(cudaMemset(dev, 0, 16 * sizeof(long long)));
  const int smem = 16384 + 65536 + 1024 + 256;
  SM3_CHECK_CUDA(cudaFuncSetAttribute(umma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  umma_rate_kernel<<<1, 288, smem>>>(n, a_from_tmem, count, ldtm_warps, dev);
  cudaError_t e = cudaDeviceSynchronize();
  if (e == cudaSuccess) e = cudaMemcpy(out_host, dev, 16 * sizeof(long long), cudaMemcpyDeviceToHost);
  cudaFree(dev);
  SM3_REQUIRE(e == cudaSuccess, SM3_ERR_CUDA, "umma_rate: %s", cudaGetErrorString(e));
  return SM3_OK;
}
