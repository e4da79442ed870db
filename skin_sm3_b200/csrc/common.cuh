// Shared helpers for the sm_100a kernels of libsm3_b200 (error state, dtype load/store, reductions).
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/sm3_b200.h"

namespace sm3 {

// ---- error state (thread local; no exceptions cross the C ABI) -------------------------------------
void set_error(const char* fmt, ...);
const char* get_error();

#define SM3_CHECK_CUDA(expr)                                                                      \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      ::sm3::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return SM3_ERR_CUDA;                                                                        \
    }                                                                                             \
  } while (0)

#define SM3_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      ::sm3::set_error(__VA_ARGS__);  \
      return (code);                  \
    }                                 \
  } while (0)

inline int dtype_size(int dt) { return dt == SM3_F32 ? 4 : 2; }
inline bool dtype_ok(int dt) { return dt == SM3_F32 || dt == SM3_F16 || dt == SM3_BF16; }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// resident CTAs per SM of a kernel (occupancy query): the grid of a persistent kernel is num_sms() * this
template <typename K>
inline int resident_ctas(K kernel, int threads, size_t dyn_smem = 0) {
  int n = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, dyn_smem) != cudaSuccess || n < 1) n = 2;
  return n;
}

// ---- programmatic dependent launch (PDL) --------------------------------------------------------------
// The fused steps enqueue 5 kernels back to back; at the reference's batch sizes each runs for a few microseconds, so the
// launch latency and prologue (barrier init, TMEM allocation, tensor-map prefetch) of kernel k+1 are worth hiding behind
// the tail of kernel k.  While a step is being enqueued `g_pdl` is set and every launch carries
// cudaLaunchAttributeProgrammaticStreamSerialization; every kernel calls pdl_wait() before its first global-memory
// access (it returns once the preceding kernel has completed and flushed) and pdl_launch() right after, which lets the
// next kernel's CTAs become resident as soon as all of ours have started.  Without the attribute both are no-ops.
// SM3_PDL=0 disables it.
extern thread_local bool g_pdl;
bool pdl_enabled();
struct PdlScope {
  bool prev;
  explicit PdlScope(bool on) : prev(g_pdl) { g_pdl = on && pdl_enabled(); }
  ~PdlScope() { g_pdl = prev; }
};
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  if (g_pdl) {
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
  }
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- scalar dtype conversion ------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ---- 128-bit vector access: VecIO<T>::N elements per 16-byte transaction ---------------------------
template <typename T> struct VecIO;
template <> struct VecIO<float> {
  static constexpr int N = 4;
  __device__ static __forceinline__ void load(const float* p, float (&o)[4]) {
    float4 v = __ldg(reinterpret_cast<const float4*>(p));
    o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
  }
  __device__ static __forceinline__ void store(float* p, const float (&o)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
  }
};
template <> struct VecIO<__half> {
  static constexpr int N = 8;
  __device__ static __forceinline__ void load(const __half* p, float (&o)[8]) {
    uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __half22float2(h[i]); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
  }
  __device__ static __forceinline__ void store(__half* p, const float (&o)[8]) {
    uint4 v;
    __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(o[2 * i], o[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = v;
  }
};
template <> struct VecIO<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static __forceinline__ void load(const __nv_bfloat16* p, float (&o)[8]) {
    uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) { float2 f = __bfloat1622float2(h[i]); o[2 * i] = f.x; o[2 * i + 1] = f.y; }
  }
  __device__ static __forceinline__ void store(__nv_bfloat16* p, const float (&o)[8]) {
    uint4 v;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(o[2 * i], o[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = v;
  }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum for blockDim.x <= 1024 (result valid in every thread)
__device__ __forceinline__ float block_sum(float v, float* smem32) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) smem32[warp] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  float r = (lane < nw) ? smem32[lane] : 0.f;
  r = warp_sum(r);
  return r;
}

// global row index of local row l (see sm3_b200.h, K2) and its positive partner
__device__ __forceinline__ int global_row(int l, int n_local, int pair_offset, int n_global) {
  return l < n_local ? pair_offset + l : n_global + pair_offset + (l - n_local);
}
__device__ __forceinline__ int positive_of(int g, int n_global) {
  return g < n_global ? g + n_global : g - n_global;
}

// ---------------------------------------------------------------------------------------------------------------
// Symmetric forward (single rank, rows == columns; infonce_tc_fwdsym_kernel): only the column tiles J >= 2R of every
// 256-row pair R are visited, and a tile with J > 2R + 1 contributes its ROW sums to rows R and its COLUMN sums to the
// rows of tile J.  The (R, J) work items form one flat list -- row pairs in the folded order 0, P-1, 1, P-2, ... so that
// every folded pair holds T + 2 tiles -- cut into pieces of `tpc` tiles, one per CTA.  Workspace layout (floats):
//   [maxseg][M]  row sums of the k-th CTA that worked on the row pair (k < nseg(R))
//   [P][M]       column sums: slab R' holds what row pair R' contributed to each later row (valid for R' < R(row))
// The folds recognise the layout by kSymFlag in n_partials (low bits = tpc); everything is summed in a fixed order.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kSymFlag = 0x40000000;
__host__ __device__ inline long sym_flat_start(int R, int T, int P) {
  const int i = R < P - 1 - R ? R : P - 1 - R;
  long f = (long)i * (T + 2);
  if (R != i) f += T - 2 * i;
  return f;
}
__host__ __device__ inline int sym_maxseg(int T, int tpc) { return (T - 1) / tpc + 2; }
__device__ __forceinline__ float fold_row_partials(const float* __restrict__ partial, int n_partials, int64_t rows,
                                                   int64_t i) {
  float s = 0.f;
  if (!(n_partials & kSymFlag)) {
    float s1 = 0.f;                               // two chains: the loads of both are in flight together
    int k = 0;
    for (; k + 2 <= n_partials; k += 2) {
      s += partial[(int64_t)k * rows + i];
      s1 += partial[(int64_t)(k + 1) * rows + i];
    }
    if (k < n_partials) s += partial[(int64_t)k * rows + i];
    return s + s1;
  }
  const int tpc = n_partials & (kSymFlag - 1);
  const int T = (int)(rows / 128), P = T / 2, R = (int)(i / 256);
  const long f = sym_flat_start(R, T, P);
  const int nseg = (int)((f + (T - 2 * R) - 1) / tpc - f / tpc) + 1;
  // one list of nseg + R addends (row-sum slabs of this row pair, then the column slabs of the earlier row pairs), eight
  // independent chains: eight strided loads in flight per round, fixed association -> deterministic
  const float* col = partial + (int64_t)sym_maxseg(T, tpc) * rows + i;
  const float* row = partial + i;
  const int total = nseg + R;
  auto term = [&](int j) { return j < nseg ? row[(int64_t)j * rows] : col[(int64_t)(j - nseg) * rows]; };
  float c[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int j = 0;
  for (; j + 8 <= total; j += 8) {
#pragma unroll
    for (int u = 0; u < 8; ++u) c[u] += term(j + u);
  }
#pragma unroll
  for (int u = 0; u < 8; ++u)
    if (j + u < total) c[u] += term(j + u);
  return ((c[0] + c[1]) + (c[2] + c[3])) + ((c[4] + c[5]) + (c[6] + c[7]));
}

// ---------------------------------------------------------------------------------------------------------------
// Symmetric forward ACROSS ranks (sm3_infonce_step_peer mode 4; infonce_tc_fwdsym_mr_kernel).  S is symmetric, so of the
// two blocks (rows of rank r, columns of rank q) and (rows of q, columns of r) only one has to be computed: its row sums
// belong to the row owner and its COLUMN sums to the column owner.  Circulant assignment: rank r computes
//   * its own block (a single-rank problem of n_local pairs in local coordinates: upper-triangular tiles, as above),
//   * the full blocks against the H = (W - 1) / 2 ranks r + 1 ... r + H (mod W),
//   * for even W half of the block against the antipodal rank r + W / 2: ranks r < W / 2 take that rank's first
//     Ja = 2 (P_l / 2) column tiles for all their row pairs (anti = 1), ranks r >= W / 2 take all of its column tiles for
//     their row pairs R >= P_l / 2, i.e. their rows from Ja * 128 on (anti = 2),
// i.e. W / 2 blocks instead of W, the same on every rank.  Per row pair R the tiles are visited own block first, then
// partner by partner (the order their rows arrive); the flat list over R is cut into pieces of tpc tiles per CTA.
// Workspace (floats, M_l = 2 n_local): [maxseg][M_l] row sums | [P_l][M_l] own-block column sums | [np][P_l][M_l] column
// sums for the partners' rows.  A second small kernel folds the partner slabs and stores one [M_l] vector per partner
// into the partner's statistics buffer (plane 2, slot = source rank); the loss kernel adds what it received.
// ---------------------------------------------------------------------------------------------------------------
struct MrPlan {
  int on, W, rank, H, anti, np;      // np = H + (anti != 0)
  int T_l, P_l, tpc, maxseg, nctas;
  long flat;                         // tiles in this rank's flat list
};
__host__ __device__ inline int mr_count(const MrPlan& m, int R) {
  int c = (m.T_l - 2 * R) + m.H * m.T_l;
  if (m.anti == 1) c += 2 * (m.P_l / 2);
  else if (m.anti == 2 && R >= m.P_l / 2) c += m.T_l;
  return c;
}
__host__ __device__ inline long mr_prefix(const MrPlan& m, int R) {      // tiles of the row pairs before R
  long f = (long)R * m.T_l - (long)R * (R - 1) + (long)m.H * m.T_l * R;
  if (m.anti == 1) f += (long)R * (2 * (m.P_l / 2));
  else if (m.anti == 2 && R > m.P_l / 2) f += (long)(R - m.P_l / 2) * m.T_l;
  return f;
}
// which rank's rows partner slot ps (0 .. np - 1) reads, and the global 128-column tile of its local tile j_l
__host__ __device__ inline int mr_partner_rank(const MrPlan& m, int ps) {
  return ps < m.H ? (m.rank + 1 + ps) % m.W : (m.rank + m.W / 2) % m.W;
}
__host__ __device__ inline int mr_global_tile(const MrPlan& m, int owner, int j_l) {
  const int half = m.T_l / 2;
  return j_l < half ? owner * half + j_l : m.W * half + owner * half + (j_l - half);
}
__device__ __forceinline__ float fold_row_partials_mr(const float* __restrict__ partial, const MrPlan& m, int64_t i) {
  const int64_t rows = (int64_t)m.T_l * 128;
  const int R = (int)(i / 256);
  const long f = mr_prefix(m, R);
  const int nseg = (int)((f + mr_count(m, R) - 1) / m.tpc - f / m.tpc) + 1;
  float s = 0.f;
  for (int k = 0; k < nseg; ++k) s += partial[(int64_t)k * rows + i];
  const float* col = partial + (int64_t)m.maxseg * rows + i;
  float c0 = 0.f, c1 = 0.f;
  int r = 0;
  for (; r + 2 <= R; r += 2) { c0 += col[(int64_t)r * rows]; c1 += col[(int64_t)(r + 1) * rows]; }
  if (r < R) c0 += col[(int64_t)r * rows];
  return s + (c0 + c1);
}

// Cross-rank flag wait used INSIDE kernels (fused exchange): one thread spins until every peer has published `epoch`
// on `channel` of this rank's flag buffer (slot layout: flags[channel * 16 + source_rank], epochs only grow).  The
// acquire at system scope orders this thread's later reads after the peers' stores.  A peer may legitimately be late
// by a long time (dataloader respawn at an epoch boundary, rank-0 checkpoint or evaluation, cudnn autotune), so the
// wait is bounded in WALL time by `timeout_ns` (host: SM3_PEER_TIMEOUT_S, default 1800 s -- the order of NCCL's own
// watchdog) and backs off with nanosleep; only past that bound does it trap, so a dead peer cannot hang the GPU for ever.
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void peer_flags_wait_all(const unsigned* flags, int world, int channel, unsigned epoch,
                                                    unsigned long long timeout_ns) {
  unsigned long long t0 = 0;
  unsigned spins = 0;
  for (int r = 0; r < world; ++r) {
    const unsigned* src = flags + channel * 16 + r;
    unsigned v;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
      if ((int)(v - epoch) >= 0) break;
      if (++spins > 64u) {                     // the first polls are back to back (the common case: data lands within us)
        __nanosleep(spins > 4096u ? 2000u : 100u);
        const unsigned long long now = globaltimer_ns();
        if (t0 == 0) t0 = now;
        if (now - t0 > timeout_ns) {
          printf("sm3: in-kernel peer wait timed out after %llu s (channel %d, peer %d, have %u, want %u); "
                 "raise SM3_PEER_TIMEOUT_S if ranks may enter the step further apart\n",
                 timeout_ns / 1000000000ull, channel, r, v, epoch);
          __trap();
        }
      }
    } while (true);
  }
}
// host side: SM3_PEER_TIMEOUT_S (seconds, default 1800; read once)
unsigned long long peer_timeout_ns();

// dtype dispatch helper for host code
#define SM3_DISPATCH_DTYPE(dt, T, ...)                       \
  do {                                                       \
    if ((dt) == SM3_F32) { using T = float; __VA_ARGS__; }   \
    else if ((dt) == SM3_F16) { using T = __half; __VA_ARGS__; } \
    else { using T = __nv_bfloat16; __VA_ARGS__; }           \
  } while (0)

// ---- kernel launch entry points implemented in the .cu files (host side) ---------------------------
int l2norm_fwd_launch(const void* p_a, int64_t rows_a, const void* p_b, int64_t rows_b, int D, int p_dtype, void* z,
                      int z_dtype, float* inv_norm, float eps, cudaStream_t st);
int l2norm_bwd_launch(const float* dz_partials, int n_partials, float scale, const void* z, int z_dtype,
                      const float* inv_norm, float eps, void* dp_a, int64_t rows_a, void* dp_b, int64_t rows_b, int D,
                      int dp_dtype, cudaStream_t st);

struct PeerPtrs {           // destination buffers of a peer scatter (one per rank, this rank included)
  void* p[16];
  int world;
};
struct InfoNceProblem {
  const void* z_rows;
  const void* z_cols;
  int n_local, pair_offset, n_global, D, dtype;
  float inv_T;
  int col_stride = 1;   // backward: element stride of the per-column statistic arrays (4 = packed float4 rows)
  int skip_local = 0;   // 1: visit only the column tiles owned by OTHER ranks (the local block is computed separately,
                        //    overlapped with the cross-rank exchange); needs n_local % 128 == 0
  const float* extra_neg_sum = nullptr;   // forward, skip_local: neg_sum of the local block, added in the finalize
  // fused exchange (tcgen05 path): the kernel itself waits for the peers' flags before it touches the first column tile
  // another rank owns (local tiles are visited first), instead of a separate wait kernel in front of it
  const unsigned* wait_flags = nullptr;
  int wait_world = 0, wait_channel = 0;
  unsigned wait_epoch = 0;
  const float* acol_direct = nullptr;     // backward: a_j = g_lse_j / neg_sum_j already materialised per global column
                                          // (written by the owners over NVLink) -> no prep kernel; g_pos_cols likewise
  int no_finalize = 0;                    // forward: leave the per-split partial sums in the workspace (the fused
                                          // loss kernel folds them)
  // forward, multi-rank, 256-row tcgen05 kernel: column tiles visited owner by owner (own columns first, then the ranks
  // whose rows arrive first) with one flag wait per source rank; when push_src is set the kernel itself stores this
  // rank's rows into the peers' column buffers (two extra warps), see infonce_tc.cu
  int push_mode = 0;
  const void* push_src = nullptr;         // this rank's normalised rows, bf16 [2 n_local, D]
  const struct PeerFused* push = nullptr; // destinations, flag buffers, [world] ticket words, rank, channel, epoch
};
struct PeerFused {          // what the fused exchange kernels need to publish to every rank
  PeerPtrs data;            // destination buffers (z_cols or stats), one per rank
  PeerPtrs flags;           // flag buffers, one per rank
  unsigned* counter;        // local zero-initialised word: CTA ticket for the "last block signals" pattern
  int rank, channel;
  unsigned epoch;
};
int l2norm_scatter_launch(const void* p_a, const void* p_b, int n_local, int pair_offset, int n_global, int D, int p_dtype,
                          void* z_local, float* inv_norm, float eps, const PeerFused& pf, cudaStream_t st);
// mr (optional): the forward was the cross-rank symmetric kernel -- fold its workspace, wait for the column sums the
// contributing ranks stored into `colsum_in` (this rank's statistics buffer, plane 2: [source rank][M_l]) and add them
struct MrFold {
  MrPlan plan;
  const float* colsum_in;
  const unsigned* flags;       // this rank's flag buffer; contributors signal channel 2
  unsigned epoch;
  unsigned long long timeout_ns;
};
int loss_stats_scatter_launch(const float* partial, int n_partials, const float* pos, int n_local, int pair_offset,
                              int n_global, float inv_T, float scale, float* loss, float* g_pos, float* g_lse,
                              float* neg_sum, float* block_ws, const PeerFused& pf, cudaStream_t st, int accumulate = 0,
                              float* a_local = nullptr, const MrFold* mr = nullptr);
// folds the partner column-sum slabs of the cross-rank symmetric forward and stores one vector per partner into the
// partner's statistics buffer (plane 2, slot = this rank), then releases flag channel 2 on the partners
int colsum_push_launch(const float* partner_slabs, const MrPlan& plan, int n_local, const PeerPtrs& stats,
                       const PeerPtrs& flags, unsigned* ticket, unsigned epoch, cudaStream_t st);
MrPlan infonce_tc_mr_plan(int n_local, int world, int rank);
size_t infonce_tc_mr_workspace(const MrPlan& m);
size_t infonce_tc_mr_workspace_any(int n_local, int world);   // max over the ranks' plans (0 when mode 4 does not apply)
int infonce_tc_fwd_mr(const InfoNceProblem& pb, const MrPlan& m, float* pos, void* ws, size_t ws_bytes, cudaStream_t st);
// byte offset of the a_j array (padded to a multiple of 64 columns + 64) inside the tcgen05 backward workspace
size_t infonce_tc_acol_offset(const InfoNceProblem& pb);
size_t infonce_simt_workspace(const InfoNceProblem& pb, int backward);
int infonce_simt_fwd(const InfoNceProblem& pb, float* pos, float* lse_neg, float* neg_sum, void* ws, size_t ws_bytes,
                     cudaStream_t st);
int infonce_simt_bwd(const InfoNceProblem& pb, const float* gpos_r, const float* glse_r, const float* nsum_r,
                     const float* gpos_c, const float* glse_c, const float* nsum_c, void* ws, size_t ws_bytes,
                     cudaStream_t st);
bool infonce_tc_supported(const InfoNceProblem& pb);
bool infonce_tc_push_supported(const InfoNceProblem& pb);   // push_mode needs the 256-row forward kernel
size_t infonce_tc_workspace(const InfoNceProblem& pb, int backward);
int infonce_tc_fwd(const InfoNceProblem& pb, float* pos, float* lse_neg, float* neg_sum, void* ws, size_t ws_bytes,
                   cudaStream_t st);
int infonce_tc_bwd(const InfoNceProblem& pb, const float* gpos_r, const float* glse_r, const float* nsum_r,
                   const float* gpos_c, const float* glse_c, const float* nsum_c, void* ws, size_t ws_bytes,
                   cudaStream_t st);
int infonce_finalize_launch(const float* partial_sums, int n_partials, int64_t rows, float inv_T, float* neg_sum,
                            float* lse_neg, cudaStream_t st, const float* extra = nullptr);
int peer_signal_launch(const PeerPtrs& flags, int rank, int channel, unsigned epoch, cudaStream_t st);
int peer_wait_launch(const unsigned* my_flags, int world, int channel, unsigned epoch, cudaStream_t st);
int peer_scatter_rows_launch(const void* src, int n_local, int pair_offset, int n_global, int row_bytes,
                             const PeerPtrs& peers, cudaStream_t st);
int peer_scatter_stats_launch(const float* g_pos, const float* g_lse, const float* nsum, int n_local, int pair_offset,
                              int n_global, const PeerPtrs& peers, cudaStream_t st);
int infonce_loss_launch(const float* pos, const float* lse_neg, int64_t rows, float scale, float* loss, int accumulate,
                        float* g_pos, float* g_lse, cudaStream_t st);

}  // namespace sm3
