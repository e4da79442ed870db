// Peer-memory exchange for the row-sharded objective: instead of an NCCL all-gather, every rank
// stores its rows straight into every rank's copy of the global buffer over NVLink/NVSwitch (plain st.global on
// peer-mapped pointers obtained from torch symmetric memory).  Two payloads use it:
//   * the normalised bf16 embedding rows  -> z_cols[2*n_global, D]      (forward)
//   * the per-row statistics (g_pos, g_lse, neg_sum) packed as float4  -> stats[2*n_global, 4]   (backward)
// Local row l lands at global row g(l) = pair_offset + l (l < n_local) or n_global + pair_offset + l - n_local, i.e.
// the reference's [all first views ; all second views] order on the concatenated batch (simclr.py:293,296-297).
// A cross-rank barrier (symmetric-memory signal pads, issued from Python on the same stream) separates the stores
// from the kernels that read the assembled buffers.
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
__global__ void peer_wait_kernel(const unsigned* __restrict__ my_flags, int world, int channel, unsigned epoch) {
  const int r = threadIdx.x;
  if (r >= world) return;
  const unsigned* src = my_flags + channel * 16 + r;
  const long long t0 = clock64();
  unsigned v;
  do {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
    if ((int)(v - epoch) >= 0) break;
    if (clock64() - t0 > 20000000000LL) {      // ~10 s: a missing peer must fail loudly, not hang the GPU
      printf("sm3: peer wait timeout (channel %d, peer %d, have %u, want %u)\n", channel, r, v, epoch);
      __trap();
    }
  } while (true);
}

}  // namespace

int peer_signal_launch(const PeerPtrs& flags, int rank, int channel, unsigned epoch, cudaStream_t st) {
  peer_signal_kernel<<<1, 32, 0, st>>>(flags, rank, channel, epoch);
  SM3_CHECK_CUDA(cudaGetLastError());
  return SM3_OK;
}
int peer_wait_launch(const unsigned* my_flags, int world, int channel, unsigned epoch, cudaStream_t st) {
  peer_wait_kernel<<<1, 32, 0, st>>>(my_flags, world, channel, epoch);
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
