// Peer-memory exchange for the row-sharded objective: instead of an NCCL all-gather, every rank
// stores its rows straight into every rank's copy of the global buffer over NVLink/NVSwitch (plain st.global on
// peer-mapped pointers obtained from torch symmetric memory).  Two payloads use it:
//   * the normalised bf16 embedding rows  -> z_cols[2*n_global, D]      (forward)
//   * the per-row statistics (g_pos, g_lse, neg_sum) packed as float4  -> stats[2*n_global, 4]   (backward)
// Local row l lands at global row g(l) = pair_offset + l (l < n_local) or n_global + pair_offset + l - n_local, i.e.
// the reference's [all first views ; all second views] order on the concatenated batch (simclr.py:293,296-297).
// A cross-rank barrier (symmetric-memory signal pads, issued from Python on the same stream) separates the stores
// from the kernels that read the assembled buffers.
#include <stdlib.h>

#include "common.cuh"

namespace sm3 {
namespace {

__global__ void __launch_bounds__(256)
peer_scatter_rows_kernel(const uint4* __restrict__ src, int n_local, int pair_offset, int n_global, int vec_per_row,
                         PeerPtrs peers) {
  const int64_t total = (int64_t)2 * n_local * vec_per_row;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int l = (int)(i / vec_per_row);
    const int v = (int)(i - (int64_t)l * vec_per_row);
    const int g = global_row(l, n_local, pair_offset, n_global);
    const uint4 val = __ldg(src + i);
    const int64_t off = (int64_t)g * vec_per_row + v;
#pragma unroll 4
    for (int r = 0; r < peers.world; ++r) reinterpret_cast<uint4*>(peers.p[r])[off] = val;
  }
}

// NVSwitch multicast variant: ONE multimem store per 16-byte vector lands in every rank's buffer (the switch
// replicates), so each GPU sends its rows once instead of `world` times.
__device__ __forceinline__ void multimem_st_v4(void* mc_addr, const uint4& v) {
  asm volatile("multimem.st.weak.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}
__global__ void __launch_bounds__(256)
peer_multicast_rows_kernel(const uint4* __restrict__ src, int n_local, int pair_offset, int n_global, int vec_per_row,
                           uint4* __restrict__ mc_dst) {
  const int64_t total = (int64_t)2 * n_local * vec_per_row;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int l = (int)(i / vec_per_row);
    const int v = (int)(i - (int64_t)l * vec_per_row);
    const int g = global_row(l, n_local, pair_offset, n_global);
    multimem_st_v4(mc_dst + (int64_t)g * vec_per_row + v, __ldg(src + i));
  }
}
__global__ void __launch_bounds__(256)
peer_multicast_stats_kernel(const float* __restrict__ g_pos, const float* __restrict__ g_lse,
                            const float* __restrict__ nsum, int n_local, int pair_offset, int n_global,
                            uint4* __restrict__ mc_dst) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= 2 * n_local) return;
  const int g = global_row(l, n_local, pair_offset, n_global);
  uint4 v;
  v.x = __float_as_uint(g_pos[l]); v.y = __float_as_uint(g_lse[l]); v.z = __float_as_uint(nsum[l]); v.w = 0u;
  multimem_st_v4(mc_dst + g, v);
}

__global__ void __launch_bounds__(256)
peer_scatter_stats_kernel(const float* __restrict__ g_pos, const float* __restrict__ g_lse,
                          const float* __restrict__ nsum, int n_local, int pair_offset, int n_global, PeerPtrs peers) {
  const int l = blockIdx.x * blockDim.x + threadIdx.x;
  if (l >= 2 * n_local) return;
  const int g = global_row(l, n_local, pair_offset, n_global);
  const float4 val = make_float4(g_pos[l], g_lse[l], nsum[l], 0.f);
  for (int r = 0; r < peers.world; ++r) reinterpret_cast<float4*>(peers.p[r])[g] = val;
}

// Split cross-rank barrier on a small symmetric flag buffer (one 32-bit slot per (channel, source rank), epochs only
// grow): `signal` publishes this rank's earlier stores to every peer, `wait` blocks the stream until every peer has
// signalled the same epoch.  Work that does not need remote data is enqueued between the two.
__global__ void peer_signal_kernel(PeerPtrs flags, int rank, int channel, unsigned epoch) {
  const int r = threadIdx.x;
  if (r >= flags.world) return;
  __threadfence_system();
  unsigned* dst = reinterpret_cast<unsigned*>(flags.p[r]) + channel * 16 + rank;
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst), "r"(epoch) : "memory");
}
__global__ void peer_wait_kernel(const unsigned* __restrict__ my_flags, int world, int channel, unsigned epoch,
                                 unsigned long long timeout_ns) {
  if (threadIdx.x == 0) peer_flags_wait_all(my_flags, world, channel, epoch, timeout_ns);   // bounded, see common.cuh
}

// ---------------------------------------------------------------------------------------------------------------
// Fused exchange kernels (SM3_PEER_FUSED=1, sm3_infonce_step_peer mode 2).  The producing kernel itself publishes to
// every rank and signals, the consuming tcgen05 kernel itself waits (peer_flags_wait_all in its prologue path), so a
// multi-rank step is 5 launches instead of 13 and the cross-rank latency hides behind the local column block.
// "Last block signals": every CTA fences its peer stores at system scope and takes a ticket; the CTA that draws the
// last ticket knows all stores of the grid are visible, resets the ticket word and releases the epoch to every peer.
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void last_block_signal(const PeerFused& pf, bool* smem_flag) {
  __syncthreads();                                   // every thread's peer stores precede thread 0's fence
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned total = gridDim.x * gridDim.y;
    const unsigned t = atomicAdd(pf.counter, 1u);
    *smem_flag = (t == total - 1);
  }
  __syncthreads();
  // Every CTA's stores were made visible system-wide by its thread 0's fence BEFORE its ticket; the CTA that sees the
  // last ticket therefore runs strictly after all of them, and so does each of its lanes below (ordered behind thread
  // 0's atomic by the barrier above): lane r publishes the epoch to rank r with its own fence + release store.
  if (*smem_flag && threadIdx.x < pf.flags.world) {
    if (threadIdx.x == 0) *pf.counter = 0u;          // next launch starts from zero (same stream => ordered)
    __threadfence_system();
    unsigned* dst = reinterpret_cast<unsigned*>(pf.flags.p[threadIdx.x]) + pf.channel * 16 + pf.rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst), "r"(pf.epoch) : "memory");
  }
}

// K1 forward + scatter: normalise this rank's rows (first halves from pa, second halves from pb), keep the bf16 rows
// locally (A operand of K2/K3, normalise-backward) and store them into every rank's z_cols at the global row index.
// One warp per 4 rows, D % 8 == 0, D <= 256 (lane < D/8 owns one 16-byte vector of the row).
template <typename TIn>
__global__ void __launch_bounds__(256)
l2norm_scatter_kernel(const TIn* __restrict__ pa, const TIn* __restrict__ pb, int n_local, int pair_offset, int n_global,
                      int D, __nv_bfloat16* __restrict__ z_local, float* __restrict__ inv_norm, float eps, PeerFused pf) {
  pdl_wait();        // PDL: the preceding kernel of the fused step has completed (no-op otherwise)
  pdl_launch();
  __shared__ bool is_last;
  constexpr int kR = 4;
  const int lane = threadIdx.x & 31;
  const int rows = 2 * n_local;
  const bool have = lane < D / 8;
  const int vpr = D / 8;
  const int row0 = (blockIdx.x * 8 + (threadIdx.x >> 5)) * kR;
  float v[kR][8];
#pragma unroll
  for (int r = 0; r < kR; ++r) {
    const int row = row0 + r;
    if (row < rows && have) {
      const TIn* src = row < n_local ? pa + (size_t)row * D : pb + (size_t)(row - n_local) * D;
      if constexpr (sizeof(TIn) == 4) {
        float a[4], b[4];
        VecIO<float>::load(reinterpret_cast<const float*>(src) + lane * 8, a);
        VecIO<float>::load(reinterpret_cast<const float*>(src) + lane * 8 + 4, b);
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[r][i] = a[i]; v[r][4 + i] = b[i]; }
      } else {
        VecIO<TIn>::load(src + lane * 8, v[r]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) v[r][i] = 0.f;
    }
  }
#pragma unroll
  for (int r = 0; r < kR; ++r) {
    const int row = row0 + r;
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) ss = fmaf(v[r][i], v[r][i], ss);
    ss = warp_sum(ss);
    const float inv = 1.0f / fmaxf(sqrtf(ss), eps);
    if (row < rows) {
      if (have) {
        uint4 pk;
        __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&pk);
#pragma unroll
        for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[r][2 * i] * inv, v[r][2 * i + 1] * inv);
        reinterpret_cast<uint4*>(z_local)[(size_t)row * vpr + lane] = pk;
        const size_t off = (size_t)global_row(row, n_local, pair_offset, n_global) * vpr + lane;
#pragma unroll 4
        for (int q = 0; q < pf.data.world; ++q) reinterpret_cast<uint4*>(pf.data.p[q])[off] = pk;
      }
      if (lane == 0) inv_norm[row] = inv;
    }
  }
  last_block_signal(pf, &is_last);
}

// 128-thread CTAs: at the launch-bound sizes (8192 rows) that is 64 CTAs instead of 32 (64 threads: no better) -- the kernel is a chain of
// dependent L2 round trips (fold, ticket, last-CTA fold), so more CTAs in flight is what shortens it
constexpr int kLossThreads = 128;

// CE on the sufficient statistics + its gradient + the backward exchange, one launch:
//   neg_sum_i = sum_k partial[k][i] ; lse_i = inv_T + log(neg_sum_i) ; loss = scale * sum_i softplus(lse_i - pos_i)
//   g_lse_i = scale * sigmoid(lse_i - pos_i) = -g_pos_i ; a_i = g_lse_i / neg_sum_i
// (g_pos, g_lse, neg_sum) stay local for this rank's rows of K3; (a_i, g_pos_i) go to every rank's column planes
// stats[0 .. M_g) = a, stats[M_g .. 2 M_g) = g_pos at the global row index.  The loss is folded by the last CTA in a
// fixed order (deterministic), which then signals.  With pf.data.world == pf.flags.world == 0 nothing is published:
// that is the single-GPU "finalize + loss" kernel (one multi-CTA launch instead of a finalize launch and a 1-CTA
// reduction over all rows).
__global__ void __launch_bounds__(kLossThreads)
loss_stats_scatter_kernel(const float* __restrict__ partial, int n_partials, const float* __restrict__ pos, int n_local,
                          int pair_offset, int n_global, float inv_T, float scale, float* __restrict__ loss,
                          float* __restrict__ g_pos, float* __restrict__ g_lse, float* __restrict__ neg_sum,
                          float* __restrict__ block_ws, PeerFused pf, int accumulate, float* __restrict__ a_local,
                          MrFold mr) {
  pdl_wait();        // PDL: the preceding kernel of the fused step has completed (no-op otherwise)
  pdl_launch();
  __shared__ float red[32];
  __shared__ bool is_last;
  const int rows = 2 * n_local;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (mr.plan.on) {
    // cross-rank symmetric forward: the column sums of the blocks other ranks computed for us arrive in plane 2 of this
    // rank's statistics buffer, one [rows] vector per contributing rank; wait for their channel-2 flags (one thread)
    if (threadIdx.x == 0) {
      const MrPlan& m = mr.plan;
      for (int c = 0; c < m.np; ++c) {
        const int src = c < m.H ? (m.rank - 1 - c + 2 * m.W) % m.W : (m.rank + m.W / 2) % m.W;
        peer_flags_wait_all(mr.flags + src, 1, 2, mr.epoch, mr.timeout_ns);
      }
    }
    __syncthreads();
  }
  float term = 0.f;
  if (i < rows) {
    float s;
    if (mr.plan.on) {
      const MrPlan& m = mr.plan;
      s = fold_row_partials_mr(partial, m, i);
      for (int c = 0; c < m.np; ++c) {                 // fixed order: rank-1, rank-2, ..., antipodal
        const int src = c < m.H ? (m.rank - 1 - c + 2 * m.W) % m.W : (m.rank + m.W / 2) % m.W;
        s += __ldcg(mr.colsum_in + (size_t)src * rows + i);
      }
    } else {
      s = fold_row_partials(partial, n_partials, rows, i);
    }
    const float lse = inv_T + logf(s);               // s == 0 (a single pair: no negatives) -> -inf, term = 0
    const float x = lse - pos[i];
    const float e = __expf(-fabsf(x));
    const float r = __frcp_rn(1.0f + e);
    const float g = scale * (x >= 0.f ? r : e * r);  // scale * sigmoid(x)
    term = fmaxf(x, 0.f) + log1pf(e);
    g_lse[i] = g;
    g_pos[i] = -g;
    neg_sum[i] = s;
    const float a = s > 0.f ? g / s : 0.f;
    if (a_local != nullptr) a_local[i] = a;          // single GPU: the backward's a_j (no separate prep launch)
    const size_t gr = (size_t)global_row(i, n_local, pair_offset, n_global);
    const size_t m_cols = (size_t)2 * n_global;
#pragma unroll 4
    for (int q = 0; q < pf.data.world; ++q) {
      float* st = reinterpret_cast<float*>(pf.data.p[q]);
      st[gr] = a;
      st[m_cols + gr] = -g;
    }
  }
  const float bs = block_sum(term, red);
  if (threadIdx.x == 0) block_ws[blockIdx.x] = bs;
  // ticket + signal; the CTA that draws the last ticket also folds the block sums
  __syncthreads();
  if (threadIdx.x == 0) {
    if (pf.data.world > 0 || pf.flags.world > 0) __threadfence_system();   // peers read what this CTA stored
    else __threadfence();                                                  // single GPU: only the last CTA reads block_ws
    const unsigned t = atomicAdd(pf.counter, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    float s = 0.f;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) s += __ldcg(block_ws + b);
    s = block_sum(s, red);
    if (threadIdx.x == 0) { *loss = (accumulate ? *loss : 0.f) + s * scale; *pf.counter = 0u; }
    if (threadIdx.x < pf.flags.world) {
      __threadfence_system();
      unsigned* dst = reinterpret_cast<unsigned*>(pf.flags.p[threadIdx.x]) + pf.channel * 16 + pf.rank;
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst), "r"(pf.epoch) : "memory");
    }
  }
}


// Cross-rank symmetric forward: column sums this rank computed for its partners' rows.  Partner slot ps holds [P_l][M_l]
// slabs (one per local row pair); the vector sum over the row pairs that visited the partner (all of them, or only the
// row pairs R >= P_l / 2 for anti == 2; only the partner's first 256 (P_l / 2) rows for anti == 1) goes to plane 2 of the partner's
// statistics buffer at [this rank][row]; the last CTA releases channel 2 on every partner.
__global__ void __launch_bounds__(256)
colsum_push_kernel(const float* __restrict__ slabs, MrPlan m, int n_local, PeerPtrs stats, PeerPtrs flags,
                   unsigned* __restrict__ ticket, unsigned epoch) {
  pdl_wait();
  pdl_launch();
  __shared__ bool is_last;
  const int64_t rows = 2 * (int64_t)n_local;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int ps = blockIdx.y;
  if (i < rows) {
    const bool anti = ps >= m.H;
    const int r_begin = (anti && m.anti == 2) ? m.P_l / 2 : 0;
    const bool covered = !(anti && m.anti == 1) || i < (int64_t)256 * (m.P_l / 2);
    float c0 = 0.f, c1 = 0.f;
    if (covered) {
      const float* col = slabs + ((int64_t)ps * m.P_l) * rows + i;
      int r = r_begin;
      for (; r + 2 <= m.P_l; r += 2) { c0 += col[(int64_t)r * rows]; c1 += col[(int64_t)(r + 1) * rows]; }
      if (r < m.P_l) c0 += col[(int64_t)r * rows];
    }
    const int dst = mr_partner_rank(m, ps);
    float* out = reinterpret_cast<float*>(stats.p[dst]) + (size_t)4 * n_local * m.W /* planes 0, 1 */ +
                 (size_t)m.rank * rows + i;
    *out = c0 + c1;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x * gridDim.y - 1);
  }
  __syncthreads();
  if (is_last) {
    if (threadIdx.x == 0) *ticket = 0u;
    if ((int)threadIdx.x < m.np) {
      __threadfence_system();
      unsigned* dst = reinterpret_cast<unsigned*>(flags.p[mr_partner_rank(m, (int)threadIdx.x)]) + 2 * 16 + m.rank;
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst), "r"(epoch) : "memory");
    }
  }
}

}  // namespace

unsigned long long peer_timeout_ns() {
  static unsigned long long ns = 0;
  if (ns == 0) {
    const char* e = getenv("SM3_PEER_TIMEOUT_S");
    double sec = e ? atof(e) : 1800.0;
    if (!(sec > 0.0)) sec = 1800.0;
    ns = (unsigned long long)(sec * 1e9);
  }
  return ns;
}

int l2norm_scatter_launch(const void* p_a, const void* p_b, int n_local, int pair_offset, int n_global, int D, int p_dtype,
                          void* z_local, float* inv_norm, float eps, const PeerFused& pf, cudaStream_t st) {
  SM3_REQUIRE(D % 8 == 0 && D >= 8 && D <= 256, SM3_ERR_SHAPE, "l2norm_scatter: D=%d", D);
  SM3_REQUIRE(aligned16(p_a) && aligned16(p_b) && aligned16(z_local), SM3_ERR_SHAPE, "l2norm_scatter: unaligned rows");
  const unsigned grid = (unsigned)((2 * (int64_t)n_local + 31) / 32);
  SM3_DISPATCH_DTYPE(p_dtype, TIn, {
    launch_k(l2norm_scatter_kernel<TIn>, dim3(grid), dim3(256), 0, st, (const TIn*)p_a, (const TIn*)p_b, n_local, pair_offset,
             n_global, D, (__nv_bfloat16*)z_local, inv_norm, eps, pf);
  });
  SM3_CHECK_CUDA(cudaGetLastError());
  return SM3_OK;
}

int loss_stats_scatter_launch(const float* partial, int n_partials, const float* pos, int n_local, int pair_offset,
                              int n_global, float inv_T, float scale, float* loss, float* g_pos, float* g_lse,
                              float* neg_sum, float* block_ws, const PeerFused& pf, cudaStream_t st, int accumulate,
                              float* a_local, const MrFold* mr) {
  const unsigned grid = (unsigned)((2 * (int64_t)n_local + kLossThreads - 1) / kLossThreads);
  MrFold mf{};
  if (mr != nullptr) mf = *mr;
  launch_k(loss_stats_scatter_kernel, dim3(grid), dim3(kLossThreads), 0, st, partial, n_partials, pos, n_local, pair_offset, n_global,
           inv_T, scale, loss, g_pos, g_lse, neg_sum, block_ws, pf, accumulate, a_local, mf);
  SM3_CHECK_CUDA(cudaGetLastError());
  return SM3_OK;
}

int colsum_push_launch(const float* partner_slabs, const MrPlan& plan, int n_local, const PeerPtrs& stats,
                       const PeerPtrs& flags, unsigned* ticket, unsigned epoch, cudaStream_t st) {
  if (plan.np < 1) return SM3_OK;
  const dim3 grid((unsigned)((2 * (int64_t)n_local + 255) / 256), (unsigned)plan.np);
  launch_k(colsum_push_kernel, grid, dim3(256), 0, st, partner_slabs, plan, n_local, stats, flags, ticket, epoch);
  SM3_CHECK_CUDA(cudaGetLastError());
  return SM3_OK;
}

int peer_signal_launch(const PeerPtrs& flags, int rank, int channel, unsigned epoch, cudaStream_t st) {
  peer_signal_kernel<<<1, 32, 0, st>>>(flags, rank, channel, epoch);
  SM3_CHECK_CUDA(cudaGetLastError());
  return SM3_OK;
}
int peer_wait_launch(const unsigned* my_flags, int world, int channel, unsigned epoch, cudaStream_t st) {
  peer_wait_kernel<<<1, 32, 0, st>>>(my_flags, world, channel, epoch, peer_timeout_ns());
  SM3_CHECK_CUDA(cudaGetLastError());
  return SM3_OK;
}

int peer_scatter_rows_launch(const void* src, int n_local, int pair_offset, int n_global, int row_bytes,
                             const PeerPtrs& peers, cudaStream_t st) {
  const int vpr = row_bytes / 16;
  const int64_t total = (int64_t)2 * n_local * vpr;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  peer_scatter_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>((const uint4*)src, n_local, pair_offset, n_global, vpr, peers);
  SM3_CHECK_CUDA(cudaGetLastError());
  return SM3_OK;
}

int peer_scatter_stats_launch(const float* g_pos, const float* g_lse, const float* nsum, int n_local, int pair_offset,
                              int n_global, const PeerPtrs& peers, cudaStream_t st) {
  peer_scatter_stats_kernel<<<(2 * n_local + 255) / 256, 256, 0, st>>>(g_pos, g_lse, nsum, n_local, pair_offset, n_global,
                                                                     peers);
  SM3_CHECK_CUDA(cudaGetLastError());
  return SM3_OK;
}

}  // namespace sm3

using namespace sm3;

extern "C" int sm3_peer_multicast_rows(const void* src, int n_local, int pair_offset, int n_global, int row_bytes,
                                       void* multicast_dst, void* stream) {
  SM3_REQUIRE(src && aligned16(src) && multicast_dst && aligned16(multicast_dst), SM3_ERR_SHAPE, "peer_multicast_rows: bad pointer");
  SM3_REQUIRE(row_bytes > 0 && row_bytes % 16 == 0, SM3_ERR_SHAPE, "peer_multicast_rows: row_bytes %d not a multiple of 16", row_bytes);
  SM3_REQUIRE(n_local >= 1 && pair_offset >= 0 && pair_offset + n_local <= n_global, SM3_ERR_SHAPE, "peer_multicast_rows: bad row block");
  const int vpr = row_bytes / 16;
  const int64_t total = (int64_t)2 * n_local * vpr;
  int64_t blocks = (total + 255) / 256;
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  peer_multicast_rows_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>((const uint4*)src, n_local, pair_offset,
                                                                                n_global, vpr, (uint4*)multicast_dst);
  SM3_CHECK_CUDA(cudaGetLastError());
  return SM3_OK;
}

extern "C" int sm3_peer_multicast_stats(const float* g_pos, const float* g_lse, const float* neg_sum, int n_local,
                                        int pair_offset, int n_global, void* multicast_dst, void* stream) {
  SM3_REQUIRE(g_pos && g_lse && neg_sum && multicast_dst && aligned16(multicast_dst), SM3_ERR_SHAPE, "peer_multicast_stats: bad pointer");
  SM3_REQUIRE(n_local >= 1 && pair_offset >= 0 && pair_offset + n_local <= n_global, SM3_ERR_SHAPE, "peer_multicast_stats: bad row block");
  peer_multicast_stats_kernel<<<(2 * n_local + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
      g_pos, g_lse, neg_sum, n_local, pair_offset, n_global, (uint4*)multicast_dst);
  SM3_CHECK_CUDA(cudaGetLastError());
  return SM3_OK;
}

static int fill_peers(PeerPtrs& pp, void* const* peers_host, int world) {
  SM3_REQUIRE(peers_host != nullptr && world >= 1 && world <= 16, SM3_ERR_SHAPE, "peer scatter: world=%d not in [1,16]", world);
  pp.world = world;
  for (int r = 0; r < world; ++r) {
    SM3_REQUIRE(peers_host[r] != nullptr && aligned16(peers_host[r]), SM3_ERR_SHAPE, "peer scatter: bad peer pointer %d", r);
    pp.p[r] = peers_host[r];
  }
  return SM3_OK;
}

extern "C" int sm3_peer_signal(void* const* peer_flags_host, int world, int rank, int channel, unsigned epoch,
                               void* stream) {
  SM3_REQUIRE(channel >= 0 && channel < 4 && rank >= 0 && rank < world, SM3_ERR_SHAPE, "peer_signal: bad channel/rank");
  PeerPtrs pp;
  int rc = fill_peers(pp, peer_flags_host, world);
  if (rc) return rc;
  return peer_signal_launch(pp, rank, channel, epoch, (cudaStream_t)stream);
}

extern "C" int sm3_peer_wait(const void* my_flags, int world, int channel, unsigned epoch, void* stream) {
  SM3_REQUIRE(my_flags && world >= 1 && world <= 16 && channel >= 0 && channel < 4, SM3_ERR_SHAPE, "peer_wait: bad argument");
  return peer_wait_launch((const unsigned*)my_flags, world, channel, epoch, (cudaStream_t)stream);
}

extern "C" int sm3_peer_scatter_rows(const void* src, int n_local, int pair_offset, int n_global, int row_bytes,
                                     void* const* peers_host, int world, void* stream) {
  SM3_REQUIRE(src && aligned16(src), SM3_ERR_SHAPE, "peer_scatter_rows: bad source pointer");
  SM3_REQUIRE(row_bytes > 0 && row_bytes % 16 == 0, SM3_ERR_SHAPE, "peer_scatter_rows: row_bytes %d not a multiple of 16", row_bytes);
  SM3_REQUIRE(n_local >= 1 && pair_offset >= 0 && pair_offset + n_local <= n_global, SM3_ERR_SHAPE, "peer_scatter_rows: bad row block");
  PeerPtrs pp;
  int rc = fill_peers(pp, peers_host, world);
  if (rc) return rc;
  return peer_scatter_rows_launch(src, n_local, pair_offset, n_global, row_bytes, pp, (cudaStream_t)stream);
}

extern "C" int sm3_peer_scatter_stats(const float* g_pos, const float* g_lse, const float* neg_sum, int n_local,
                                      int pair_offset, int n_global, void* const* peers_host, int world, void* stream) {
  SM3_REQUIRE(g_pos && g_lse && neg_sum, SM3_ERR_SHAPE, "peer_scatter_stats: null pointer");
  SM3_REQUIRE(n_local >= 1 && pair_offset >= 0 && pair_offset + n_local <= n_global, SM3_ERR_SHAPE, "peer_scatter_stats: bad row block");
  PeerPtrs pp;
  int rc = fill_peers(pp, peers_host, world);
  if (rc) return rc;
  return peer_scatter_stats_launch(g_pos, g_lse, neg_sum, n_local, pair_offset, n_global, pp, (cudaStream_t)stream);
}
