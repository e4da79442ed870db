// extern "C" surface of libsm3_b200.so (see include/sm3_b200.h).  Argument validation, algorithm choice
// (tcgen05 vs fp32-FMA kernels -- both CUDA, there is no CPU path), and the host-buffer convenience entry.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "common.cuh"

#include <mutex>
#include <utility>
#include <vector>

namespace sm3 {

static thread_local char g_err[512] = "";
thread_local bool g_pdl = false;
bool pdl_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("SM3_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

// ---- stage timing (bench / profiling only): events recorded between the stages of the fused steps ----
static constexpr int kMaxStages = 8;
static bool g_stage_timing = false;
static cudaEvent_t g_stage_ev[kMaxStages + 1] = {};
static int g_stage_count = 0;                 // events recorded by the most recent step (stages = count - 1)
static const char* g_stage_names = "";
static void stage_begin(cudaStream_t st, const char* names) {
  if (!g_stage_timing) return;
  g_stage_names = names;
  g_stage_count = 0;
  for (int i = 0; i <= kMaxStages; ++i)
    if (g_stage_ev[i] == nullptr && cudaEventCreate(&g_stage_ev[i]) != cudaSuccess) { g_stage_timing = false; return; }
  cudaEventRecord(g_stage_ev[g_stage_count++], st);
}
static void stage_mark(cudaStream_t st) {
  if (g_stage_timing && g_stage_count > 0 && g_stage_count <= kMaxStages) cudaEventRecord(g_stage_ev[g_stage_count++], st);
}

static int pick_algo(const InfoNceProblem& pb, int algo) {
  if (algo == SM3_ALGO_AUTO) return infonce_tc_supported(pb) ? SM3_ALGO_TC : SM3_ALGO_SIMT;
  return algo;
}

static int check_problem(const InfoNceProblem& pb) {
  SM3_REQUIRE(pb.z_rows && pb.z_cols, SM3_ERR_SHAPE, "infonce: null embedding pointer");
  SM3_REQUIRE(dtype_ok(pb.dtype), SM3_ERR_DTYPE, "infonce: bad dtype %d", pb.dtype);
  SM3_REQUIRE(pb.inv_T > 0.f && pb.inv_T * 2.0f * 1.4426950f < 240.f, SM3_ERR_SHAPE,
              "infonce: temperature %g outside the constant-shift range (need 1/T < 83)", 1.0 / pb.inv_T);
  return SM3_OK;
}

}  // namespace sm3

using namespace sm3;

extern "C" int sm3_version(void) { return SM3_ABI_VERSION; }

extern "C" int sm3_stage_timing(int enable) {
  g_stage_timing = enable != 0;
  g_stage_count = 0;
  return SM3_OK;
}
extern "C" const char* sm3_stage_timing_names(void) { return g_stage_names; }
extern "C" int sm3_stage_timing_read(float* ms, int capacity) {
  SM3_REQUIRE(ms != nullptr && capacity >= 1, SM3_ERR_SHAPE, "stage_timing_read: bad buffer");
  if (g_stage_count < 2) return 0;
  SM3_CHECK_CUDA(cudaEventSynchronize(g_stage_ev[g_stage_count - 1]));
  int n = 0;
  for (int i = 1; i < g_stage_count && n < capacity; ++i, ++n)
    SM3_CHECK_CUDA(cudaEventElapsedTime(&ms[n], g_stage_ev[i - 1], g_stage_ev[i]));
  return n;
}
extern "C" const char* sm3_last_error(void) { return get_error(); }

extern "C" int sm3_device_supported(void) {
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    set_error("no usable CUDA device");
    return SM3_ERR_CUDA;
  }
  return prop.major == 10 ? 1 : 0;
}

extern "C" int sm3_l2norm_fwd(const void* p_a, int64_t rows_a, const void* p_b, int64_t rows_b, int D, int p_dtype,
                              void* z, int z_dtype, float* inv_norm, float eps, void* stream) {
  SM3_REQUIRE(p_a && z && inv_norm, SM3_ERR_SHAPE, "l2norm_fwd: null pointer");
  SM3_REQUIRE(rows_a >= 0 && rows_b >= 0 && (rows_b == 0 || p_b), SM3_ERR_SHAPE, "l2norm_fwd: bad row counts");
  SM3_REQUIRE(D >= 1, SM3_ERR_SHAPE, "l2norm_fwd: D=%d", D);
  SM3_REQUIRE(dtype_ok(p_dtype) && dtype_ok(z_dtype), SM3_ERR_DTYPE, "l2norm_fwd: bad dtype");
  SM3_REQUIRE(eps > 0.f, SM3_ERR_SHAPE, "l2norm_fwd: eps must be > 0");
  return l2norm_fwd_launch(p_a, rows_a, p_b, rows_b, D, p_dtype, z, z_dtype, inv_norm, eps, (cudaStream_t)stream);
}

extern "C" int sm3_l2norm_bwd(const float* dz_partials, int n_partials, float scale, const void* z, int z_dtype,
                              const float* inv_norm, float eps, void* dp_a, int64_t rows_a, void* dp_b,
                              int64_t rows_b, int D, int dp_dtype, void* stream) {
  SM3_REQUIRE(dz_partials && z && inv_norm && dp_a, SM3_ERR_SHAPE, "l2norm_bwd: null pointer");
  SM3_REQUIRE(n_partials >= 1, SM3_ERR_SHAPE, "l2norm_bwd: n_partials=%d", n_partials);
  SM3_REQUIRE(rows_a >= 0 && rows_b >= 0 && (rows_b == 0 || dp_b), SM3_ERR_SHAPE, "l2norm_bwd: bad row counts");
  SM3_REQUIRE(dtype_ok(z_dtype) && dtype_ok(dp_dtype), SM3_ERR_DTYPE, "l2norm_bwd: bad dtype");
  return l2norm_bwd_launch(dz_partials, n_partials, scale, z, z_dtype, inv_norm, eps, dp_a, rows_a, dp_b, rows_b, D,
                           dp_dtype, (cudaStream_t)stream);
}

extern "C" size_t sm3_infonce_workspace_bytes(int n_local, int n_global, int D, int dtype, int algo, int backward) {
  InfoNceProblem pb{nullptr, nullptr, n_local, 0, n_global, D, dtype, 1.0f};
  // AUTO must cover whichever kernel the call ends up using
  size_t a = 0, b = 0;
  if (algo != SM3_ALGO_TC) a = infonce_simt_workspace(pb, backward);
  if (algo != SM3_ALGO_SIMT) b = infonce_tc_workspace(pb, backward);
  return a > b ? a : b;
}

extern "C" int sm3_infonce_fwd(const void* z_rows, const void* z_cols, int n_local, int pair_offset, int n_global,
                               int D, int dtype, float inv_T, float* pos, float* lse_neg, float* neg_sum,
                               void* workspace, size_t workspace_bytes, int algo, void* stream) {
  InfoNceProblem pb{z_rows, z_cols, n_local, pair_offset, n_global, D, dtype, inv_T};
  int rc = check_problem(pb);
  if (rc) return rc;
  SM3_REQUIRE(pos && lse_neg && neg_sum && workspace, SM3_ERR_SHAPE, "infonce_fwd: null output/workspace");
  const int a = pick_algo(pb, algo);
  if (a == SM3_ALGO_TC) {
    SM3_REQUIRE(infonce_tc_supported(pb), SM3_ERR_DTYPE,
                "infonce_fwd: tcgen05 path needs bf16 rows, D in {64,128,192,256}, 16-byte aligned (got dtype=%d D=%d)",
                dtype, D);
    return infonce_tc_fwd(pb, pos, lse_neg, neg_sum, workspace, workspace_bytes, (cudaStream_t)stream);
  }
  SM3_REQUIRE(a == SM3_ALGO_SIMT, SM3_ERR_SHAPE, "infonce_fwd: unknown algo %d", algo);
  return infonce_simt_fwd(pb, pos, lse_neg, neg_sum, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int sm3_infonce_bwd(const void* z_rows, const void* z_cols, int n_local, int pair_offset, int n_global,
                               int D, int dtype, float inv_T, const float* g_pos_rows, const float* g_lse_rows,
                               const float* neg_sum_rows, const float* g_pos_cols, const float* g_lse_cols,
                               const float* neg_sum_cols, void* workspace, size_t workspace_bytes, int algo,
                               void* stream) {
  InfoNceProblem pb{z_rows, z_cols, n_local, pair_offset, n_global, D, dtype, inv_T};
  int rc = check_problem(pb);
  if (rc) return rc;
  SM3_REQUIRE(g_pos_rows && g_lse_rows && neg_sum_rows && g_pos_cols && g_lse_cols && neg_sum_cols && workspace,
              SM3_ERR_SHAPE, "infonce_bwd: null pointer");
  const int a = pick_algo(pb, algo);
  if (a == SM3_ALGO_TC) {
    SM3_REQUIRE(infonce_tc_supported(pb), SM3_ERR_DTYPE,
                "infonce_bwd: tcgen05 path needs bf16 rows, D in {64,128,192,256}, 16-byte aligned (got dtype=%d D=%d)",
                dtype, D);
    return infonce_tc_bwd(pb, g_pos_rows, g_lse_rows, neg_sum_rows, g_pos_cols, g_lse_cols, neg_sum_cols, workspace,
                          workspace_bytes, (cudaStream_t)stream);
  }
  SM3_REQUIRE(a == SM3_ALGO_SIMT, SM3_ERR_SHAPE, "infonce_bwd: unknown algo %d", algo);
  return infonce_simt_bwd(pb, g_pos_rows, g_lse_rows, neg_sum_rows, g_pos_cols, g_lse_cols, neg_sum_cols, workspace,
                          workspace_bytes, (cudaStream_t)stream);
}

extern "C" int sm3_infonce_bwd_packed(const void* z_rows, const void* z_cols, int n_local, int pair_offset,
                                      int n_global, int D, int dtype, float inv_T, const float* g_pos_rows,
                                      const float* g_lse_rows, const float* neg_sum_rows, const float* stats_cols,
                                      void* workspace, size_t workspace_bytes, int algo, void* stream) {
  InfoNceProblem pb{z_rows, z_cols, n_local, pair_offset, n_global, D, dtype, inv_T};
  pb.col_stride = 4;
  int rc = check_problem(pb);
  if (rc) return rc;
  SM3_REQUIRE(g_pos_rows && g_lse_rows && neg_sum_rows && stats_cols && workspace, SM3_ERR_SHAPE,
              "infonce_bwd_packed: null pointer");
  const int a = pick_algo(pb, algo);
  if (a == SM3_ALGO_TC) {
    SM3_REQUIRE(infonce_tc_supported(pb), SM3_ERR_DTYPE, "infonce_bwd_packed: tcgen05 path needs bf16 rows, D in {64,128,192,256}");
    return infonce_tc_bwd(pb, g_pos_rows, g_lse_rows, neg_sum_rows, stats_cols, stats_cols + 1, stats_cols + 2, workspace,
                          workspace_bytes, (cudaStream_t)stream);
  }
  SM3_REQUIRE(a == SM3_ALGO_SIMT, SM3_ERR_SHAPE, "infonce_bwd_packed: unknown algo %d", algo);
  return infonce_simt_bwd(pb, g_pos_rows, g_lse_rows, neg_sum_rows, stats_cols, stats_cols + 1, stats_cols + 2, workspace,
                          workspace_bytes, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------------
// host-buffer entry: H2D -> normalise -> K2 -> loss -> K3 -> normalise-backward -> D2H, one stream.
// ---------------------------------------------------------------------------------------------------
namespace {
struct HostPlan {
  size_t p1, p2, z, inv, pos, lse, nsum, gpos, glse, loss, dp1, dp2, ws, ws_bytes, total;
  int z_dtype, algo;
};
size_t align_up(size_t v) { return (v + 255) & ~(size_t)255; }

HostPlan plan_host(int n, int D, int io_dtype, int algo) {
  HostPlan h{};
  const size_t m = 2 * (size_t)n;
  InfoNceProblem probe{(const void*)256, (const void*)256, n, 0, n, D, SM3_BF16, 10.f};
  int a = algo;
  if (a == SM3_ALGO_AUTO) a = infonce_tc_supported(probe) ? SM3_ALGO_TC : SM3_ALGO_SIMT;
  h.algo = a;
  h.z_dtype = (a == SM3_ALGO_TC) ? SM3_BF16 : SM3_F32;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes); return r; };
  h.p1 = take((size_t)n * D * dtype_size(io_dtype));
  h.p2 = take((size_t)n * D * dtype_size(io_dtype));
  h.z = take(m * D * dtype_size(h.z_dtype));
  h.inv = take(m * 4); h.pos = take(m * 4); h.lse = take(m * 4); h.nsum = take(m * 4);
  h.gpos = take(m * 4); h.glse = take(m * 4); h.loss = take(256);
  h.dp1 = take((size_t)n * D * dtype_size(io_dtype));
  h.dp2 = take((size_t)n * D * dtype_size(io_dtype));
  InfoNceProblem pb{nullptr, nullptr, n, 0, n, D, h.z_dtype, 1.f};
  const size_t w0 = (a == SM3_ALGO_TC) ? infonce_tc_workspace(pb, 0) : infonce_simt_workspace(pb, 0);
  const size_t w1 = (a == SM3_ALGO_TC) ? infonce_tc_workspace(pb, 1) : infonce_simt_workspace(pb, 1);
  h.ws_bytes = w0 > w1 ? w0 : w1;
  h.ws = take(h.ws_bytes);
  h.total = o;
  return h;
}
}  // namespace

// ---------------------------------------------------------------------------------------------------
// "remote" halves of the row-sharded kernels: only the column tiles owned by OTHER ranks are visited (tcgen05 path).
// The local block is an ordinary single-rank call on the local rows (n_global = n_local, pair_offset = 0); it needs
// no remote data, so it runs while the exchange is in flight.  See skin_sm3_b200/functional.py (_FusedInfoNCE).
// ---------------------------------------------------------------------------------------------------
extern "C" size_t sm3_infonce_remote_workspace_bytes(int n_local, int n_global, int D, int backward) {
  InfoNceProblem pb{nullptr, nullptr, n_local, 0, n_global, D, SM3_BF16, 1.0f};
  pb.skip_local = 1;
  return infonce_tc_workspace(pb, backward);
}

extern "C" int sm3_infonce_fwd_remote(const void* z_rows, const void* z_cols, int n_local, int pair_offset,
                                      int n_global, int D, int dtype, float inv_T, const float* neg_sum_local,
                                      float* pos_unused, float* lse_neg, float* neg_sum, void* workspace,
                                      size_t workspace_bytes, void* stream) {
  InfoNceProblem pb{z_rows, z_cols, n_local, pair_offset, n_global, D, dtype, inv_T};
  pb.skip_local = 1;
  pb.extra_neg_sum = neg_sum_local;
  int rc = check_problem(pb);
  if (rc) return rc;
  SM3_REQUIRE(neg_sum_local && pos_unused && lse_neg && neg_sum && workspace, SM3_ERR_SHAPE, "infonce_fwd_remote: null pointer");
  SM3_REQUIRE(infonce_tc_supported(pb), SM3_ERR_DTYPE, "infonce_fwd_remote: needs the tcgen05 path (bf16 rows, D in {64,128,192,256})");
  return infonce_tc_fwd(pb, pos_unused, lse_neg, neg_sum, workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int sm3_infonce_bwd_remote_packed(const void* z_rows, const void* z_cols, int n_local, int pair_offset,
                                             int n_global, int D, int dtype, float inv_T, const float* g_pos_rows,
                                             const float* g_lse_rows, const float* neg_sum_rows,
                                             const float* stats_cols, void* workspace, size_t workspace_bytes,
                                             void* stream) {
  InfoNceProblem pb{z_rows, z_cols, n_local, pair_offset, n_global, D, dtype, inv_T};
  pb.skip_local = 1;
  pb.col_stride = 4;
  int rc = check_problem(pb);
  if (rc) return rc;
  SM3_REQUIRE(g_pos_rows && g_lse_rows && neg_sum_rows && stats_cols && workspace, SM3_ERR_SHAPE, "infonce_bwd_remote: null pointer");
  SM3_REQUIRE(infonce_tc_supported(pb), SM3_ERR_DTYPE, "infonce_bwd_remote: needs the tcgen05 path (bf16 rows, D in {64,128,192,256})");
  return infonce_tc_bwd(pb, g_pos_rows, g_lse_rows, neg_sum_rows, stats_cols, stats_cols + 1, stats_cols + 2, workspace,
                        workspace_bytes, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------------
// device-pointer fused step: normalise -> K2 -> loss -> K3 -> normalise-backward enqueued by ONE call
// (no host<->device copies, no synchronisation).  Same scratch layout as the host entry.
// ---------------------------------------------------------------------------------------------------
extern "C" size_t sm3_infonce_step_scratch_bytes(int n_pairs, int D, int io_dtype, int algo) {
  if (n_pairs < 1 || D < 1 || !dtype_ok(io_dtype)) return 0;
  return plan_host(n_pairs, D, io_dtype, algo).total;
}

static int step_impl(const void* p1, const void* p2, int n_pairs, int D, int io_dtype, float temperature, float weight,
                     float* loss, int accumulate, void* dp1, void* dp2, void* device_scratch, size_t scratch_bytes,
                     int algo, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SM3_REQUIRE(p1 && p2 && loss && device_scratch, SM3_ERR_SHAPE, "infonce_step: null pointer");
  SM3_REQUIRE((dp1 == nullptr) == (dp2 == nullptr), SM3_ERR_SHAPE, "infonce_step: dp1/dp2 must both be given or both NULL");
  SM3_REQUIRE(n_pairs >= 1 && D >= 1 && dtype_ok(io_dtype), SM3_ERR_SHAPE, "infonce_step: bad shape/dtype");
  SM3_REQUIRE(temperature > 0.f, SM3_ERR_SHAPE, "infonce_step: temperature must be > 0");
  HostPlan h = plan_host(n_pairs, D, io_dtype, algo);
  if (h.algo == SM3_ALGO_TC && !(aligned16(p1) && aligned16(p2))) { h = plan_host(n_pairs, D, io_dtype, SM3_ALGO_SIMT); }
  SM3_REQUIRE(scratch_bytes >= h.total, SM3_ERR_WORKSPACE, "infonce_step: scratch %zu < %zu", scratch_bytes, h.total);
  char* base = (char*)device_scratch;
  const int64_t n = n_pairs, m = 2 * n;
  const float inv_T = 1.0f / temperature;
  unsigned* ticket = (unsigned*)(base + h.loss + 64);
  if (h.algo == SM3_ALGO_TC) SM3_CHECK_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned), st));   // ahead of the kernel chain
  PdlScope pdl(!g_stage_timing);          // event records between the kernels would break the chain anyway
  stage_begin(st, "normalize,infonce_fwd,loss,infonce_bwd,normalize_bwd");
  int rc = sm3_l2norm_fwd(p1, n, p2, n, D, io_dtype, base + h.z, h.z_dtype, (float*)(base + h.inv), 1e-12f, st);
  if (rc) return rc;
  stage_mark(st);
  if (h.algo == SM3_ALGO_TC) {
    // K2 leaves its per-split partial row sums in the workspace; ONE multi-CTA kernel folds them and does the CE on
    // the statistics + its gradient (instead of a finalize launch and a single-CTA reduction over all 2N rows)
    InfoNceProblem pb{base + h.z, base + h.z, n_pairs, 0, n_pairs, D, h.z_dtype, inv_T};
    pb.no_finalize = 1;
    SM3_REQUIRE(infonce_tc_supported(pb), SM3_ERR_DTYPE, "infonce_step: tcgen05 path unavailable for this shape");
    const int splits = infonce_tc_fwd(pb, (float*)(base + h.pos), (float*)(base + h.lse), (float*)(base + h.nsum),
                                      base + h.ws, h.ws_bytes, st);
    if (splits < 0) return splits;
    stage_mark(st);
    PeerFused none{};
    none.counter = ticket;
    // the same kernel also materialises a_j = g_lse_j / neg_sum_j where K3 expects it (behind the partial-gradient slabs:
    // K2's partial sums at the start of the workspace are dead by the time K3 overwrites them), so the backward needs no
    // prep launch
    float* acol = (float*)(base + h.ws + infonce_tc_acol_offset(pb));
    rc = loss_stats_scatter_launch((const float*)(base + h.ws), splits, (const float*)(base + h.pos), n_pairs, 0, n_pairs,
                                   inv_T, weight / (float)m, loss, (float*)(base + h.gpos), (float*)(base + h.glse),
                                   (float*)(base + h.nsum), (float*)(base + h.lse) /* per-CTA loss sums */, none, st,
                                   accumulate, dp1 ? acol : nullptr);
    stage_mark(st);
    if (rc || !dp1) return rc;
    InfoNceProblem pk{base + h.z, base + h.z, n_pairs, 0, n_pairs, D, h.z_dtype, inv_T};
    pk.acol_direct = acol;
    const int np = infonce_tc_bwd(pk, (float*)(base + h.gpos), (float*)(base + h.glse), (float*)(base + h.nsum),
                                  (float*)(base + h.gpos), (float*)(base + h.glse), (float*)(base + h.nsum), base + h.ws,
                                  h.ws_bytes, st);
    if (np < 0) return np;
    stage_mark(st);
    rc = sm3_l2norm_bwd((const float*)(base + h.ws), np, 1.0f, base + h.z, h.z_dtype, (float*)(base + h.inv), 1e-12f,
                        dp1, n, dp2, n, D, io_dtype, st);
    stage_mark(st);
    return rc;
  } else {
    rc = sm3_infonce_fwd(base + h.z, base + h.z, n_pairs, 0, n_pairs, D, h.z_dtype, inv_T, (float*)(base + h.pos),
                         (float*)(base + h.lse), (float*)(base + h.nsum), base + h.ws, h.ws_bytes, h.algo, st);
    if (rc) return rc;
    if (rc) return rc;
    stage_mark(st);
    rc = sm3_infonce_loss((float*)(base + h.pos), (float*)(base + h.lse), m, weight / (float)m, loss, accumulate,
                          dp1 ? (float*)(base + h.gpos) : nullptr, dp1 ? (float*)(base + h.glse) : nullptr, st);
    stage_mark(st);
  }
  if (rc || !dp1) return rc;
  const int np = sm3_infonce_bwd(base + h.z, base + h.z, n_pairs, 0, n_pairs, D, h.z_dtype, inv_T,
                                 (float*)(base + h.gpos), (float*)(base + h.glse), (float*)(base + h.nsum),
                                 (float*)(base + h.gpos), (float*)(base + h.glse), (float*)(base + h.nsum),
                                 base + h.ws, h.ws_bytes, h.algo, st);
  if (np < 0) return np;
  stage_mark(st);
  rc = sm3_l2norm_bwd((const float*)(base + h.ws), np, 1.0f, base + h.z, h.z_dtype, (float*)(base + h.inv), 1e-12f,
                      dp1, n, dp2, n, D, io_dtype, st);
  stage_mark(st);
  return rc;
}

extern "C" int sm3_infonce_step(const void* p1, const void* p2, int n_pairs, int D, int io_dtype, float temperature,
                                float weight, float* loss, void* dp1, void* dp2, void* device_scratch,
                                size_t scratch_bytes, int algo, void* stream) {
  return step_impl(p1, p2, n_pairs, D, io_dtype, temperature, weight, loss, 0, dp1, dp2, device_scratch, scratch_bytes,
                   algo, stream);
}

// Grouped form: `num_terms` independent InfoNCE terms of one shape (the derm / clinic / cross / cross terms of
// SimCLRSkinV3 style 0, tools/backbone_train.py:101-121) enqueued by ONE call; loss = sum_t weights[t] * CE_t in one
// device scalar.  The terms run back to back on `stream` and share the scratch of a single step.
extern "C" int sm3_infonce_step_multi(int num_terms, const void* const* p1_host_array, const void* const* p2_host_array,
                                      int n_pairs, int D, int io_dtype, float temperature, const float* weights_host,
                                      float* loss, void* const* dp1_host_array, void* const* dp2_host_array,
                                      void* device_scratch, size_t scratch_bytes, int algo, void* stream) {
  SM3_REQUIRE(num_terms >= 1 && num_terms <= 64, SM3_ERR_SHAPE, "infonce_step_multi: num_terms=%d not in [1,64]", num_terms);
  SM3_REQUIRE(p1_host_array && p2_host_array && loss && device_scratch, SM3_ERR_SHAPE, "infonce_step_multi: null pointer");
  SM3_REQUIRE((dp1_host_array == nullptr) == (dp2_host_array == nullptr), SM3_ERR_SHAPE,
              "infonce_step_multi: dp1/dp2 arrays must both be given or both NULL");
  SM3_REQUIRE(n_pairs >= 1 && D >= 1 && dtype_ok(io_dtype) && temperature > 0.f, SM3_ERR_SHAPE,
              "infonce_step_multi: bad shape/dtype/temperature");
  const size_t need = plan_host(n_pairs, D, io_dtype, algo).total;
  SM3_REQUIRE(scratch_bytes >= need, SM3_ERR_WORKSPACE, "infonce_step_multi: scratch %zu < %zu", scratch_bytes, need);
  for (int t = 0; t < num_terms; ++t) {
    SM3_REQUIRE(p1_host_array[t] && p2_host_array[t], SM3_ERR_SHAPE, "infonce_step_multi: term %d has a null input", t);
    void* d1 = dp1_host_array ? dp1_host_array[t] : nullptr;
    void* d2 = dp2_host_array ? dp2_host_array[t] : nullptr;
    SM3_REQUIRE((d1 == nullptr) == (d2 == nullptr), SM3_ERR_SHAPE, "infonce_step_multi: term %d: dp1/dp2 mismatch", t);
    const float w = weights_host ? weights_host[t] : 1.0f;
    // the loss kernel of term t > 0 adds into the scalar written by term 0 (same stream => ordered)
    const int rc = step_impl(p1_host_array[t], p2_host_array[t], n_pairs, D, io_dtype, temperature, w, loss, t > 0, d1, d2,
                             device_scratch, scratch_bytes, algo, stream);
    if (rc) return rc;
  }
  return SM3_OK;
}

extern "C" size_t sm3_infonce_step_multi_scratch_bytes(int n_pairs, int D, int io_dtype, int algo) {
  if (n_pairs < 1 || D < 1 || !dtype_ok(io_dtype)) return 0;
  return plan_host(n_pairs, D, io_dtype, algo).total;
}

// ---------------------------------------------------------------------------------------------------
// multi-rank fused step with the peer-memory exchange, enqueued by ONE call on two streams:
//   main: normalise | local K2 | wait | remote K2 | loss | local K3 | wait | remote K3 | normalise-backward
//   side:           scatter z + signal            ....        scatter stats + signal
// (at 8 ranks the per-rank GPU work is ~0.7 ms; a Python-orchestrated step of ~25 launches is CPU-bound there)
// ---------------------------------------------------------------------------------------------------
namespace {
struct PeerPlan {
  size_t z, inv, pos, lse_l, ns_l, lse, nsum, gpos, glse, ws_f, ws_f_bytes, ws_b, ws_b_bytes, total;
};
PeerPlan plan_peer(int n_local, int n_global, int D) {
  PeerPlan h{};
  const size_t m = 2 * (size_t)n_local;
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o = align_up(o + bytes); return r; };
  h.z = take(m * D * 2);
  h.inv = take(m * 4); h.pos = take(m * 4); h.lse_l = take(m * 4); h.ns_l = take(m * 4);
  h.lse = take(m * 4); h.nsum = take(m * 4); h.gpos = take(m * 4); h.glse = take(m * 4);
  InfoNceProblem loc{nullptr, nullptr, n_local, 0, n_local, D, SM3_BF16, 1.f};
  InfoNceProblem rem{nullptr, nullptr, n_local, 0, n_global, D, SM3_BF16, 1.f};
  rem.skip_local = 1;
  const size_t f0 = infonce_tc_workspace(loc, 0), f1 = infonce_tc_workspace(rem, 0);
  h.ws_f_bytes = f0 > f1 ? f0 : f1;
  h.ws_f = take(h.ws_f_bytes);
  InfoNceProblem full{nullptr, nullptr, n_local, 0, n_global, D, SM3_BF16, 1.f};
  const size_t split_b = infonce_tc_workspace(loc, 1) + infonce_tc_workspace(rem, 1) + 1024;
  const size_t full_b = infonce_tc_workspace(full, 1), full_f = infonce_tc_workspace(full, 0);
  h.ws_b_bytes = split_b > full_b ? split_b : full_b;
  if (h.ws_b_bytes < full_f) h.ws_b_bytes = full_f;
  if (n_local % 128 == 0 && n_global > n_local) {            // mode 3: the owner-ordered forward may use other split counts
    InfoNceProblem fullp = full;
    fullp.push_mode = 1;
    const size_t f3 = infonce_tc_workspace(fullp, 0);
    if (h.ws_b_bytes < f3) h.ws_b_bytes = f3;
  }
  if (n_local % 128 == 0 && n_global > n_local && n_global % n_local == 0 && n_global / n_local <= 16) {
    // mode 4: the symmetric forward across ranks keeps its row / column slabs here.  Sized for EVERY rank (the plan differs
    // slightly between the two antipodal classes), so that whether mode 4 applies never depends on the rank.
    const size_t f4 = infonce_tc_mr_workspace_any(n_local, n_global / n_local);
    if (h.ws_b_bytes < f4) h.ws_b_bytes = f4;
  }
  h.ws_b = take(h.ws_b_bytes);
  h.total = o;
  return h;
}
cudaEvent_t* peer_events() {
  static thread_local cudaEvent_t ev[2] = {nullptr, nullptr};
  if (ev[0] == nullptr) {
    if (cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming) != cudaSuccess)
      return nullptr;
  }
  return ev;
}
int fill_peer_ptrs(PeerPtrs& pp, void* const* host, int world) {
  SM3_REQUIRE(host != nullptr && world >= 2 && world <= 16, SM3_ERR_SHAPE, "infonce_step_peer: world=%d not in [2,16]", world);
  pp.world = world;
  for (int r = 0; r < world; ++r) {
    SM3_REQUIRE(host[r] != nullptr && aligned16(host[r]), SM3_ERR_SHAPE, "infonce_step_peer: bad peer pointer %d", r);
    pp.p[r] = host[r];
  }
  return SM3_OK;
}
}  // namespace

extern "C" size_t sm3_infonce_step_peer_scratch_bytes(int n_local, int n_global, int D) {
  if (n_local < 1 || n_global < n_local || D % 64 != 0 || D < 64 || D > 256) return 0;
  return plan_peer(n_local, n_global, D).total;
}

extern "C" int sm3_infonce_step_peer(const void* p1, const void* p2, int n_local, int rank, int world, int D,
                                     int io_dtype, float temperature, float weight, float* loss, void* dp1, void* dp2,
                                     void* z_cols_mine, void* const* z_peers_host, void* stats_mine,
                                     void* const* stats_peers_host, void* flags_mine, void* const* flags_peers_host,
                                     unsigned epoch, int overlap, void* device_scratch, size_t scratch_bytes,
                                     void* stream_main, void* stream_side) {
  cudaStream_t sm = (cudaStream_t)stream_main, ss = (cudaStream_t)stream_side;
  SM3_REQUIRE(p1 && p2 && loss && z_cols_mine && stats_mine && flags_mine && device_scratch, SM3_ERR_SHAPE,
              "infonce_step_peer: null pointer");
  SM3_REQUIRE((dp1 == nullptr) == (dp2 == nullptr), SM3_ERR_SHAPE, "infonce_step_peer: dp1/dp2 must both be given or both NULL");
  SM3_REQUIRE(world >= 2 && rank >= 0 && rank < world && n_local >= 1, SM3_ERR_SHAPE,
              "infonce_step_peer: needs world >= 2 (got world=%d n_local=%d)", world, n_local);
  SM3_REQUIRE(overlap >= 0 && overlap <= 4, SM3_ERR_SHAPE, "infonce_step_peer: mode %d not in {0,1,2,3,4}", overlap);
  SM3_REQUIRE(!overlap || n_local % 128 == 0, SM3_ERR_SHAPE, "infonce_step_peer: modes 1 to 4 need n_local %% 128 == 0");
  SM3_REQUIRE(D % 64 == 0 && D >= 64 && D <= 256 && dtype_ok(io_dtype) && temperature > 0.f, SM3_ERR_DTYPE,
              "infonce_step_peer: D must be in {64,128,192,256}");
  SM3_REQUIRE(overlap != 1 || stream_side != stream_main, SM3_ERR_SHAPE,
              "infonce_step_peer: the side stream must differ from the main stream");
  const int n_global = n_local * world, off = rank * n_local;
  const PeerPlan h = plan_peer(n_local, n_global, D);
  SM3_REQUIRE(scratch_bytes >= h.total, SM3_ERR_WORKSPACE, "infonce_step_peer: scratch %zu < %zu", scratch_bytes, h.total);
  PeerPtrs zp, sp, fp;
  int rc = fill_peer_ptrs(zp, z_peers_host, world);
  if (rc) return rc;
  if ((rc = fill_peer_ptrs(sp, stats_peers_host, world))) return rc;
  if ((rc = fill_peer_ptrs(fp, flags_peers_host, world))) return rc;
  cudaEvent_t* ev = peer_events();
  SM3_REQUIRE(ev != nullptr, SM3_ERR_CUDA, "infonce_step_peer: cudaEventCreate failed");
  char* base = (char*)device_scratch;
  const int64_t n = n_local, m = 2 * n;
  const float inv_T = 1.0f / temperature;
  float *pos = (float*)(base + h.pos), *lse_l = (float*)(base + h.lse_l), *ns_l = (float*)(base + h.ns_l);
  float *lse = (float*)(base + h.lse), *nsum = (float*)(base + h.nsum), *gpos = (float*)(base + h.gpos), *glse = (float*)(base + h.glse);
  void* z = base + h.z;

  if (overlap == 3) {
    // ---- fused exchange with the row push INSIDE K2 (5 launches): kernel 1 only normalises (local rows + this rank's
    //      own rows of its column buffer); K2's two extra warps push the rows to the peers, destination rank+1 first,
    //      while its tensor-core warps work through this rank's own column tiles, and it waits per source rank in the
    //      order the rows arrive (rank-1, rank-2, ...).  Falls back to mode 2 where the 256-row kernel is not used. ----
    InfoNceProblem probe{base + h.z, z_cols_mine, n_local, off, n_global, D, SM3_BF16, inv_T};
    if (!infonce_tc_push_supported(probe)) overlap = 2;
  }
  if (overlap == 3) {
    SM3_REQUIRE(aligned16(p1) && aligned16(p2), SM3_ERR_SHAPE, "infonce_step_peer: fused mode needs 16-byte aligned rows");
    unsigned* counters = (unsigned*)flags_mine + 64;          // [0,2) mode-2 tickets | [2,18) push tickets | [20] kernel 1
    PeerFused own{};
    own.data.world = 1; own.data.p[0] = z_cols_mine;          // only this rank's copy of the column buffer, no signal
    own.flags.world = 0; own.counter = counters + 20; own.rank = rank; own.channel = 0; own.epoch = epoch;
    PdlScope pdl(!g_stage_timing);
    stage_begin(sm, "normalize,infonce_fwd_push,loss_scatter,infonce_bwd,normalize_bwd");
    rc = l2norm_scatter_launch(p1, p2, n_local, off, n_global, D, io_dtype, z, (float*)(base + h.inv), 1e-12f, own, sm);
    if (rc) return rc;
    stage_mark(sm);
    PeerFused pz{zp, fp, counters + 2, rank, 0, epoch};
    InfoNceProblem pf{z, z_cols_mine, n_local, off, n_global, D, SM3_BF16, inv_T};
    pf.wait_flags = (const unsigned*)flags_mine; pf.wait_world = world; pf.wait_channel = 0; pf.wait_epoch = epoch;
    pf.no_finalize = 1;
    pf.push_mode = 1; pf.push_src = z; pf.push = &pz;
    const int splits = infonce_tc_fwd(pf, pos, lse, nsum, base + h.ws_b, h.ws_b_bytes, sm);
    if (splits < 0) return splits;
    stage_mark(sm);
    PeerFused ps{sp, fp, counters + 1, rank, 1, epoch};
    rc = loss_stats_scatter_launch((const float*)(base + h.ws_b), splits, pos, n_local, off, n_global, inv_T,
                                   weight / (float)m, loss, gpos, glse, nsum, lse_l /* per-CTA loss sums */, ps, sm);
    stage_mark(sm);
    if (rc || !dp1) return rc;
    InfoNceProblem pk{z, z_cols_mine, n_local, off, n_global, D, SM3_BF16, inv_T};
    pk.wait_flags = (const unsigned*)flags_mine; pk.wait_world = world; pk.wait_channel = 1; pk.wait_epoch = epoch;
    pk.acol_direct = (const float*)stats_mine;                 // planes written by the owners: a_j | g_pos_j
    const float* gpos_cols = (const float*)stats_mine + (size_t)2 * n_global;
    const int np = infonce_tc_bwd(pk, gpos, glse, nsum, gpos_cols, gpos_cols, gpos_cols, base + h.ws_b, h.ws_b_bytes, sm);
    if (np < 0) return np;
    stage_mark(sm);
    rc = sm3_l2norm_bwd((const float*)(base + h.ws_b), np, 1.0f, z, SM3_BF16, (float*)(base + h.inv), 1e-12f, dp1, n,
                        dp2, n, D, io_dtype, sm);
    stage_mark(sm);
    return rc;
  }
  if (overlap == 4) {
    // ---- fused exchange + SYMMETRIC forward across ranks (6 launches): every rank computes W / 2 of the W column blocks
    //      of its row block (own block upper-triangular, the next (W-1)/2 ranks' blocks, half of the antipodal one) and
    //      sends the column sums of the foreign blocks to their owners (see MrPlan, common.cuh).  Falls back to mode 2
    //      when the plan does not apply (n_local %% 128, workspace). ----
    const MrPlan mp = infonce_tc_mr_plan(n_local, world, rank);
    if (!mp.on || infonce_tc_mr_workspace_any(n_local, world) > h.ws_b_bytes) overlap = 2;   // rank-independent decision
  }
  if (overlap == 4) {
    SM3_REQUIRE(aligned16(p1) && aligned16(p2), SM3_ERR_SHAPE, "infonce_step_peer: fused mode needs 16-byte aligned rows");
    const MrPlan mp = infonce_tc_mr_plan(n_local, world, rank);
    unsigned* counters = (unsigned*)flags_mine + 64;          // [0,2) tickets of modes 2 / 4 | [24] column-sum push
    PeerFused pz{zp, fp, counters, rank, 0, epoch};
    PdlScope pdl(!g_stage_timing);
    stage_begin(sm, "normalize_scatter,infonce_fwd_sym,colsum_push,loss_scatter,infonce_bwd,normalize_bwd");
    rc = l2norm_scatter_launch(p1, p2, n_local, off, n_global, D, io_dtype, z, (float*)(base + h.inv), 1e-12f, pz, sm);
    if (rc) return rc;
    stage_mark(sm);
    InfoNceProblem pf{z, z_cols_mine, n_local, off, n_global, D, SM3_BF16, inv_T};
    pf.wait_flags = (const unsigned*)flags_mine; pf.wait_world = world; pf.wait_channel = 0; pf.wait_epoch = epoch;
    SM3_REQUIRE(infonce_tc_supported(pf), SM3_ERR_DTYPE, "infonce_step_peer: tcgen05 path unavailable");
    rc = infonce_tc_fwd_mr(pf, mp, pos, base + h.ws_b, h.ws_b_bytes, sm);
    if (rc) return rc;
    stage_mark(sm);
    const float* slabs = (const float*)(base + h.ws_b) + (size_t)(mp.maxseg + mp.P_l) * (size_t)m;
    rc = colsum_push_launch(slabs, mp, n_local, sp, fp, counters + 24, epoch, sm);
    if (rc) return rc;
    stage_mark(sm);
    PeerFused ps{sp, fp, counters + 1, rank, 1, epoch};
    MrFold mf{mp, (const float*)stats_mine + (size_t)4 * n_global, (const unsigned*)flags_mine, epoch, peer_timeout_ns()};
    rc = loss_stats_scatter_launch((const float*)(base + h.ws_b), 0, pos, n_local, off, n_global, inv_T,
                                   weight / (float)m, loss, gpos, glse, nsum, lse_l /* per-CTA loss sums */, ps, sm, 0,
                                   nullptr, &mf);
    stage_mark(sm);
    if (rc || !dp1) return rc;
    InfoNceProblem pk{z, z_cols_mine, n_local, off, n_global, D, SM3_BF16, inv_T};
    pk.wait_flags = (const unsigned*)flags_mine; pk.wait_world = world; pk.wait_channel = 1; pk.wait_epoch = epoch;
    pk.acol_direct = (const float*)stats_mine;                 // planes written by the owners: a_j | g_pos_j
    const float* gpos_cols = (const float*)stats_mine + (size_t)2 * n_global;
    const int np = infonce_tc_bwd(pk, gpos, glse, nsum, gpos_cols, gpos_cols, gpos_cols, base + h.ws_b, h.ws_b_bytes, sm);
    if (np < 0) return np;
    stage_mark(sm);
    rc = sm3_l2norm_bwd((const float*)(base + h.ws_b), np, 1.0f, z, SM3_BF16, (float*)(base + h.inv), 1e-12f, dp1, n,
                        dp2, n, D, io_dtype, sm);
    stage_mark(sm);
    return rc;
  }
  if (overlap == 2) {
    // ---- fused exchange: 5 launches.  The producers publish + signal themselves, K2 / K3 wait inside the kernel and
    //      visit this rank's own column tiles first (see peer.cu, infonce_tc.cu). ----
    SM3_REQUIRE(aligned16(p1) && aligned16(p2), SM3_ERR_SHAPE, "infonce_step_peer: fused mode needs 16-byte aligned rows");
    unsigned* counters = (unsigned*)flags_mine + 64;          // two local ticket words behind the 64 flag slots
    PeerFused pz{zp, fp, counters, rank, 0, epoch};
    PdlScope pdl(!g_stage_timing);
    stage_begin(sm, "normalize_scatter,infonce_fwd,loss_scatter,infonce_bwd,normalize_bwd");
    rc = l2norm_scatter_launch(p1, p2, n_local, off, n_global, D, io_dtype, z, (float*)(base + h.inv), 1e-12f, pz, sm);
    if (rc) return rc;
    stage_mark(sm);
    InfoNceProblem pf{z, z_cols_mine, n_local, off, n_global, D, SM3_BF16, inv_T};
    pf.wait_flags = (const unsigned*)flags_mine; pf.wait_world = world; pf.wait_channel = 0; pf.wait_epoch = epoch;
    pf.no_finalize = 1;
    SM3_REQUIRE(infonce_tc_supported(pf), SM3_ERR_DTYPE, "infonce_step_peer: tcgen05 path unavailable");
    const int splits = infonce_tc_fwd(pf, pos, lse, nsum, base + h.ws_b, h.ws_b_bytes, sm);
    if (splits < 0) return splits;
    stage_mark(sm);
    PeerFused ps{sp, fp, counters + 1, rank, 1, epoch};
    rc = loss_stats_scatter_launch((const float*)(base + h.ws_b), splits, pos, n_local, off, n_global, inv_T,
                                   weight / (float)m, loss, gpos, glse, nsum, lse_l /* per-CTA loss sums */, ps, sm);
    stage_mark(sm);
    if (rc || !dp1) return rc;
    InfoNceProblem pk{z, z_cols_mine, n_local, off, n_global, D, SM3_BF16, inv_T};
    pk.wait_flags = (const unsigned*)flags_mine; pk.wait_world = world; pk.wait_channel = 1; pk.wait_epoch = epoch;
    pk.acol_direct = (const float*)stats_mine;                 // planes written by the owners: a_j | g_pos_j
    const float* gpos_cols = (const float*)stats_mine + (size_t)2 * n_global;
    const int np = infonce_tc_bwd(pk, gpos, glse, nsum, gpos_cols, gpos_cols, gpos_cols, base + h.ws_b, h.ws_b_bytes, sm);
    if (np < 0) return np;
    stage_mark(sm);
    rc = sm3_l2norm_bwd((const float*)(base + h.ws_b), np, 1.0f, z, SM3_BF16, (float*)(base + h.inv), 1e-12f, dp1, n,
                        dp2, n, D, io_dtype, sm);
    stage_mark(sm);
    return rc;
  }
  rc = sm3_l2norm_fwd(p1, n, p2, n, D, io_dtype, z, SM3_BF16, (float*)(base + h.inv), 1e-12f, sm);
  if (rc) return rc;
  if (!overlap) {
    // ---- everything on the main stream: exchange, barrier, full-width kernels (best below ~8 ranks) ----
    if ((rc = peer_scatter_rows_launch(z, n_local, off, n_global, D * 2, zp, sm))) return rc;
    if ((rc = peer_signal_launch(fp, rank, 0, epoch, sm))) return rc;
    if ((rc = peer_wait_launch((const unsigned*)flags_mine, world, 0, epoch, sm))) return rc;
    rc = sm3_infonce_fwd(z, z_cols_mine, n_local, off, n_global, D, SM3_BF16, inv_T, pos, lse, nsum, base + h.ws_b,
                         h.ws_b_bytes, SM3_ALGO_TC, sm);
    if (rc) return rc;
    rc = sm3_infonce_loss(pos, lse, m, weight / (float)m, loss, 0, dp1 ? gpos : nullptr, dp1 ? glse : nullptr, sm);
    if (rc || !dp1) return rc;
    if ((rc = peer_scatter_stats_launch(gpos, glse, nsum, n_local, off, n_global, sp, sm))) return rc;
    if ((rc = peer_signal_launch(fp, rank, 1, epoch, sm))) return rc;
    if ((rc = peer_wait_launch((const unsigned*)flags_mine, world, 1, epoch, sm))) return rc;
    const int np = sm3_infonce_bwd_packed(z, z_cols_mine, n_local, off, n_global, D, SM3_BF16, inv_T, gpos, glse, nsum,
                                          (const float*)stats_mine, base + h.ws_b, h.ws_b_bytes, SM3_ALGO_TC, sm);
    if (np < 0) return np;
    return sm3_l2norm_bwd((const float*)(base + h.ws_b), np, 1.0f, z, SM3_BF16, (float*)(base + h.inv), 1e-12f, dp1, n,
                          dp2, n, D, io_dtype, sm);
  }
  // ---- exchange of the normalised rows on the side stream, local column block meanwhile ----
  SM3_CHECK_CUDA(cudaEventRecord(ev[0], sm));
  SM3_CHECK_CUDA(cudaStreamWaitEvent(ss, ev[0], 0));
  if ((rc = peer_scatter_rows_launch(z, n_local, off, n_global, D * 2, zp, ss))) return rc;
  if ((rc = peer_signal_launch(fp, rank, 0, epoch, ss))) return rc;
  SM3_CHECK_CUDA(cudaEventRecord(ev[1], ss));
  rc = sm3_infonce_fwd(z, z, n_local, 0, n_local, D, SM3_BF16, inv_T, pos, lse_l, ns_l, base + h.ws_f, h.ws_f_bytes,
                       SM3_ALGO_TC, sm);
  if (rc) return rc;
  SM3_CHECK_CUDA(cudaStreamWaitEvent(sm, ev[1], 0));
  if ((rc = peer_wait_launch((const unsigned*)flags_mine, world, 0, epoch, sm))) return rc;
  rc = sm3_infonce_fwd_remote(z, z_cols_mine, n_local, off, n_global, D, SM3_BF16, inv_T, ns_l, pos, lse, nsum,
                              base + h.ws_f, h.ws_f_bytes, sm);
  if (rc) return rc;
  rc = sm3_infonce_loss(pos, lse, m, weight / (float)m, loss, 0, dp1 ? gpos : nullptr, dp1 ? glse : nullptr, sm);
  if (rc || !dp1) return rc;
  // ---- exchange of the 12-byte row statistics, local column block of the backward meanwhile ----
  SM3_CHECK_CUDA(cudaEventRecord(ev[0], sm));
  SM3_CHECK_CUDA(cudaStreamWaitEvent(ss, ev[0], 0));
  if ((rc = peer_scatter_stats_launch(gpos, glse, nsum, n_local, off, n_global, sp, ss))) return rc;
  if ((rc = peer_signal_launch(fp, rank, 1, epoch, ss))) return rc;
  SM3_CHECK_CUDA(cudaEventRecord(ev[1], ss));
  InfoNceProblem loc{nullptr, nullptr, n_local, 0, n_local, D, SM3_BF16, 1.f};
  const size_t wsl = infonce_tc_workspace(loc, 1);
  const int np_l = sm3_infonce_bwd(z, z, n_local, 0, n_local, D, SM3_BF16, inv_T, gpos, glse, nsum, gpos, glse, nsum,
                                   base + h.ws_b, wsl, SM3_ALGO_TC, sm);
  if (np_l < 0) return np_l;
  SM3_CHECK_CUDA(cudaStreamWaitEvent(sm, ev[1], 0));
  if ((rc = peer_wait_launch((const unsigned*)flags_mine, world, 1, epoch, sm))) return rc;
  const size_t slab = (size_t)np_l * m * D * 4;
  const int np_r = sm3_infonce_bwd_remote_packed(z, z_cols_mine, n_local, off, n_global, D, SM3_BF16, inv_T, gpos, glse,
                                                 nsum, (const float*)stats_mine, base + h.ws_b + slab,
                                                 h.ws_b_bytes - slab, sm);
  if (np_r < 0) return np_r;
  return sm3_l2norm_bwd((const float*)(base + h.ws_b), np_l + np_r, 1.0f, z, SM3_BF16, (float*)(base + h.inv), 1e-12f,
                        dp1, n, dp2, n, D, io_dtype, sm);
}

// debug / single-GPU test of the owner-ordered forward (mode 3's tile mapping and per-source flag waits) without peers:
// `flags` must already hold `epoch` in slots [0, world) of channel 0 (the caller plays the part of the other ranks and
// has filled z_cols); no rows are pushed.  Same outputs as sm3_infonce_fwd.
extern "C" int sm3_debug_infonce_fwd_ordered(const void* z_rows, const void* z_cols, int n_local, int rank, int world, int D,
                                             float inv_T, const void* flags, unsigned epoch, float* pos, float* lse_neg,
                                             float* neg_sum, void* workspace, size_t workspace_bytes, void* stream) {
  SM3_REQUIRE(z_rows && z_cols && flags && pos && lse_neg && neg_sum && workspace, SM3_ERR_SHAPE, "fwd_ordered: null pointer");
  SM3_REQUIRE(world >= 2 && world <= 16 && rank >= 0 && rank < world && n_local >= 128 && n_local % 128 == 0, SM3_ERR_SHAPE,
              "fwd_ordered: needs 2 <= world <= 16 and n_local %% 128 == 0");
  InfoNceProblem pb{z_rows, z_cols, n_local, rank * n_local, n_local * world, D, SM3_BF16, inv_T};
  int rc = check_problem(pb);
  if (rc) return rc;
  SM3_REQUIRE(infonce_tc_supported(pb), SM3_ERR_DTYPE, "fwd_ordered: tcgen05 path unavailable");
  pb.wait_flags = (const unsigned*)flags; pb.wait_world = world; pb.wait_channel = 0; pb.wait_epoch = epoch;
  pb.push_mode = 1;
  return infonce_tc_fwd(pb, pos, lse_neg, neg_sum, workspace, workspace_bytes, (cudaStream_t)stream);
}
extern "C" size_t sm3_debug_infonce_fwd_ordered_workspace(int n_local, int world, int D) {
  InfoNceProblem pb{nullptr, nullptr, n_local, 0, n_local * world, D, SM3_BF16, 1.0f};
  pb.push_mode = 1;
  return infonce_tc_workspace(pb, 0);
}

// debug / single-GPU test of mode 4 (symmetric forward across ranks) without peers: the caller plays all ranks on one
// device, one after the other.  _forward runs rank `rank`'s K2 (its flags must already hold `epoch` in channel 0 for every
// source) and the column-sum push into the given per-rank statistics / flag buffers; _fold runs rank `rank`'s loss kernel
// (nothing published) once every rank's _forward has been enqueued, and leaves neg_sum / lse_neg / loss.
extern "C" size_t sm3_debug_mr_workspace(int n_local, int world, int rank) {
  const MrPlan mp = infonce_tc_mr_plan(n_local, world, rank);
  return mp.on ? infonce_tc_mr_workspace(mp) : 0;
}
extern "C" int sm3_debug_mr_forward(const void* z_local, const void* z_all, int n_local, int world, int rank, int D,
                                    float inv_T, void* flags_mine, void* const* stats_peers_host,
                                    void* const* flags_peers_host, unsigned epoch, float* pos, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SM3_REQUIRE(z_local && z_all && flags_mine && pos && workspace, SM3_ERR_SHAPE, "debug_mr_forward: null pointer");
  const MrPlan mp = infonce_tc_mr_plan(n_local, world, rank);
  SM3_REQUIRE(mp.on, SM3_ERR_SHAPE, "debug_mr_forward: no plan for n_local=%d world=%d", n_local, world);
  PeerPtrs sp, fp;
  int rc = fill_peer_ptrs(sp, stats_peers_host, world);
  if (rc) return rc;
  if ((rc = fill_peer_ptrs(fp, flags_peers_host, world))) return rc;
  InfoNceProblem pf{z_local, z_all, n_local, rank * n_local, n_local * world, D, SM3_BF16, inv_T};
  pf.wait_flags = (const unsigned*)flags_mine; pf.wait_world = world; pf.wait_channel = 0; pf.wait_epoch = epoch;
  SM3_REQUIRE(infonce_tc_supported(pf), SM3_ERR_DTYPE, "debug_mr_forward: tcgen05 path unavailable");
  rc = infonce_tc_fwd_mr(pf, mp, pos, workspace, workspace_bytes, st);
  if (rc) return rc;
  const float* slabs = (const float*)workspace + (size_t)(mp.maxseg + mp.P_l) * (size_t)(2 * n_local);
  return colsum_push_launch(slabs, mp, n_local, sp, fp, (unsigned*)flags_mine + 64 + 24, epoch, st);
}
extern "C" int sm3_debug_mr_fold(const void* workspace, int n_local, int world, int rank, float inv_T, const float* pos,
                                 const void* stats_mine, void* flags_mine, unsigned epoch, float* loss, float* neg_sum,
                                 float* g_pos, float* g_lse, float* block_ws, void* stream) {
  SM3_REQUIRE(workspace && pos && stats_mine && flags_mine && loss && neg_sum && g_pos && g_lse && block_ws, SM3_ERR_SHAPE,
              "debug_mr_fold: null pointer");
  const MrPlan mp = infonce_tc_mr_plan(n_local, world, rank);
  SM3_REQUIRE(mp.on, SM3_ERR_SHAPE, "debug_mr_fold: no plan");
  PeerFused none{};
  none.counter = (unsigned*)flags_mine + 64 + 1;
  MrFold mf{mp, (const float*)stats_mine + (size_t)4 * n_local * world, (const unsigned*)flags_mine, epoch, peer_timeout_ns()};
  return loss_stats_scatter_launch((const float*)workspace, 0, pos, n_local, rank * n_local, n_local * world, inv_T,
                                   1.0f / (2.0f * n_local), loss, g_pos, g_lse, neg_sum, block_ws, none, (cudaStream_t)stream,
                                   0, nullptr, &mf);
}

extern "C" size_t sm3_infonce_host_scratch_bytes(int n_pairs, int D, int io_dtype, int algo) {
  if (n_pairs < 1 || D < 1 || !dtype_ok(io_dtype)) return 0;
  return plan_host(n_pairs, D, io_dtype, algo).total;
}

extern "C" int sm3_infonce_host(const void* p1_host, const void* p2_host, int n_pairs, int D, int io_dtype,
                                float temperature, float* loss_host, void* dp1_host, void* dp2_host,
                                void* device_scratch, size_t scratch_bytes, int algo, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  SM3_REQUIRE(p1_host && p2_host && loss_host && device_scratch, SM3_ERR_SHAPE, "infonce_host: null pointer");
  SM3_REQUIRE(n_pairs >= 1 && D >= 1 && dtype_ok(io_dtype), SM3_ERR_SHAPE, "infonce_host: bad shape/dtype");
  SM3_REQUIRE(temperature > 0.f, SM3_ERR_SHAPE, "infonce_host: temperature must be > 0");
  const HostPlan h = plan_host(n_pairs, D, io_dtype, algo);
  SM3_REQUIRE(scratch_bytes >= h.total, SM3_ERR_WORKSPACE, "infonce_host: scratch %zu < %zu", scratch_bytes, h.total);
  char* base = (char*)device_scratch;
  const size_t in_bytes = (size_t)n_pairs * D * dtype_size(io_dtype);
  SM3_CHECK_CUDA(cudaMemcpyAsync(base + h.p1, p1_host, in_bytes, cudaMemcpyHostToDevice, st));
  SM3_CHECK_CUDA(cudaMemcpyAsync(base + h.p2, p2_host, in_bytes, cudaMemcpyHostToDevice, st));
  const bool grads = dp1_host && dp2_host;
  const int rc = step_impl(base + h.p1, base + h.p2, n_pairs, D, io_dtype, temperature, 1.0f, (float*)(base + h.loss), 0,
                           grads ? base + h.dp1 : nullptr, grads ? base + h.dp2 : nullptr, device_scratch, scratch_bytes,
                           algo, stream);
  if (rc) return rc;
  SM3_CHECK_CUDA(cudaMemcpyAsync(loss_host, base + h.loss, 4, cudaMemcpyDeviceToHost, st));
  if (grads) {
    SM3_CHECK_CUDA(cudaMemcpyAsync(dp1_host, base + h.dp1, in_bytes, cudaMemcpyDeviceToHost, st));
    SM3_CHECK_CUDA(cudaMemcpyAsync(dp2_host, base + h.dp2, in_bytes, cudaMemcpyDeviceToHost, st));
  }
  SM3_CHECK_CUDA(cudaStreamSynchronize(st));
  return SM3_OK;
}

// ---------------------------------------------------------------------------------------------------
// pipelined host-buffer entry: three streams owned by the handle (H2D | kernels | D2H) and `depth` I/O slots, so the
// copies of step k+1 and k-1 run on the copy engines while the kernels of step k occupy the SMs.  Everything else
// (normalised rows, statistics, partial-gradient workspace) is touched by the kernel stream only and exists once.
// ---------------------------------------------------------------------------------------------------
struct sm3_host_pipe {
  int n_pairs, D, io_dtype, algo, depth, device;
  int n_global;                 // > n_pairs: peer mode (n_pairs = this rank's pairs), steps go through sm3_infonce_step_peer
  size_t in_bytes, io_stride, step_bytes, total;
  char* base;
  cudaStream_t s_h2d, s_run, s_d2h;
  cudaEvent_t ev_in[SM3_PIPE_MAX_DEPTH], ev_run[SM3_PIPE_MAX_DEPTH], ev_out[SM3_PIPE_MAX_DEPTH];
  int busy[SM3_PIPE_MAX_DEPTH];
  int64_t next_ticket;
};

namespace {
// per-slot I/O block: p1 | p2 | dp1 | dp2 | loss
size_t pipe_io_stride(size_t in_bytes) { return 4 * align_up(in_bytes) + 256; }
}  // namespace

extern "C" size_t sm3_host_pipe_scratch_bytes(int n_pairs, int D, int io_dtype, int algo, int depth) {
  if (n_pairs < 1 || D < 1 || !dtype_ok(io_dtype) || depth < 1 || depth > SM3_PIPE_MAX_DEPTH) return 0;
  const size_t in_bytes = (size_t)n_pairs * D * dtype_size(io_dtype);
  return align_up(plan_host(n_pairs, D, io_dtype, algo).total) + (size_t)depth * pipe_io_stride(in_bytes);
}

extern "C" int sm3_host_pipe_create(sm3_host_pipe** out, int n_pairs, int D, int io_dtype, int algo, int depth,
                                    void* device_scratch, size_t scratch_bytes) {
  SM3_REQUIRE(out != nullptr && device_scratch != nullptr, SM3_ERR_SHAPE, "host_pipe_create: null pointer");
  *out = nullptr;
  SM3_REQUIRE(n_pairs >= 1 && D >= 1 && dtype_ok(io_dtype), SM3_ERR_SHAPE, "host_pipe_create: bad shape/dtype");
  SM3_REQUIRE(depth >= 1 && depth <= SM3_PIPE_MAX_DEPTH, SM3_ERR_SHAPE, "host_pipe_create: depth %d not in [1,%d]", depth,
              SM3_PIPE_MAX_DEPTH);
  const size_t need = sm3_host_pipe_scratch_bytes(n_pairs, D, io_dtype, algo, depth);
  SM3_REQUIRE(scratch_bytes >= need, SM3_ERR_WORKSPACE, "host_pipe_create: scratch %zu < %zu", scratch_bytes, need);
  SM3_REQUIRE(aligned16(device_scratch), SM3_ERR_SHAPE, "host_pipe_create: scratch must be 16-byte aligned");
  sm3_host_pipe* hp = new (std::nothrow) sm3_host_pipe();
  SM3_REQUIRE(hp != nullptr, SM3_ERR_CUDA, "host_pipe_create: out of host memory");
  hp->n_pairs = n_pairs; hp->D = D; hp->io_dtype = io_dtype; hp->algo = algo; hp->depth = depth; hp->n_global = n_pairs;
  hp->in_bytes = (size_t)n_pairs * D * dtype_size(io_dtype);
  hp->io_stride = pipe_io_stride(hp->in_bytes);
  hp->step_bytes = align_up(plan_host(n_pairs, D, io_dtype, algo).total);
  hp->total = need;
  hp->base = (char*)device_scratch;
  hp->next_ticket = 0;
  hp->s_h2d = hp->s_run = hp->s_d2h = nullptr;
  for (int i = 0; i < SM3_PIPE_MAX_DEPTH; ++i) { hp->ev_in[i] = hp->ev_run[i] = hp->ev_out[i] = nullptr; hp->busy[i] = 0; }
  cudaError_t e = cudaGetDevice(&hp->device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&hp->s_h2d, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&hp->s_run, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&hp->s_d2h, cudaStreamNonBlocking);
  for (int i = 0; i < depth && e == cudaSuccess; ++i) {
    e = cudaEventCreateWithFlags(&hp->ev_in[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&hp->ev_run[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&hp->ev_out[i], cudaEventDisableTiming);
  }
  if (e != cudaSuccess) {
    set_error("host_pipe_create: %s", cudaGetErrorString(e));
    sm3_host_pipe_destroy(hp);
    return SM3_ERR_CUDA;
  }
  *out = hp;
  return SM3_OK;
}

extern "C" int sm3_host_pipe_destroy(sm3_host_pipe* hp) {
  if (hp == nullptr) return SM3_OK;
  if (hp->s_d2h) cudaStreamSynchronize(hp->s_d2h);
  if (hp->s_run) cudaStreamSynchronize(hp->s_run);
  if (hp->s_h2d) cudaStreamSynchronize(hp->s_h2d);
  for (int i = 0; i < SM3_PIPE_MAX_DEPTH; ++i) {
    if (hp->ev_in[i]) cudaEventDestroy(hp->ev_in[i]);
    if (hp->ev_run[i]) cudaEventDestroy(hp->ev_run[i]);
    if (hp->ev_out[i]) cudaEventDestroy(hp->ev_out[i]);
  }
  if (hp->s_h2d) cudaStreamDestroy(hp->s_h2d);
  if (hp->s_run) cudaStreamDestroy(hp->s_run);
  if (hp->s_d2h) cudaStreamDestroy(hp->s_d2h);
  delete hp;
  return SM3_OK;
}

extern "C" int64_t sm3_host_pipe_submit(sm3_host_pipe* hp, const void* p1_host, const void* p2_host, float temperature,
                                        float* loss_host, void* dp1_host, void* dp2_host) {
  SM3_REQUIRE(hp && p1_host && p2_host && loss_host, SM3_ERR_SHAPE, "host_pipe_submit: null pointer");
  SM3_REQUIRE((dp1_host == nullptr) == (dp2_host == nullptr), SM3_ERR_SHAPE,
              "host_pipe_submit: dp1_host/dp2_host must both be given or both NULL");
  SM3_REQUIRE(temperature > 0.f, SM3_ERR_SHAPE, "host_pipe_submit: temperature must be > 0");
  int dev = -1;
  SM3_CHECK_CUDA(cudaGetDevice(&dev));
  SM3_REQUIRE(dev == hp->device, SM3_ERR_SHAPE, "host_pipe_submit: handle belongs to device %d, current device is %d",
              hp->device, dev);
  const int64_t ticket = hp->next_ticket;
  const int s = (int)(ticket % hp->depth);
  // back-pressure: the previous occupant of this slot must have delivered its results to the host
  if (hp->busy[s]) SM3_CHECK_CUDA(cudaEventSynchronize(hp->ev_out[s]));
  char* io = hp->base + hp->step_bytes + (size_t)s * hp->io_stride;
  const size_t a = align_up(hp->in_bytes);
  char *p1 = io, *p2 = io + a, *dp1 = io + 2 * a, *dp2 = io + 3 * a;
  float* loss = (float*)(io + 4 * a);
  // H2D: the slot's input block was last read by the kernels of ticket - depth (ev_run[s])
  if (hp->busy[s]) SM3_CHECK_CUDA(cudaStreamWaitEvent(hp->s_h2d, hp->ev_run[s], 0));
  SM3_CHECK_CUDA(cudaMemcpyAsync(p1, p1_host, hp->in_bytes, cudaMemcpyHostToDevice, hp->s_h2d));
  SM3_CHECK_CUDA(cudaMemcpyAsync(p2, p2_host, hp->in_bytes, cudaMemcpyHostToDevice, hp->s_h2d));
  SM3_CHECK_CUDA(cudaEventRecord(hp->ev_in[s], hp->s_h2d));
  // kernels: wait for the inputs; the slot's output block is free because ev_out[s] was synchronised above
  SM3_CHECK_CUDA(cudaStreamWaitEvent(hp->s_run, hp->ev_in[s], 0));
  const int rc = sm3_infonce_step(p1, p2, hp->n_pairs, hp->D, hp->io_dtype, temperature, 1.0f, loss,
                                  dp1_host ? dp1 : nullptr, dp1_host ? dp2 : nullptr, hp->base, hp->step_bytes, hp->algo,
                                  hp->s_run);
  if (rc) return rc;
  SM3_CHECK_CUDA(cudaEventRecord(hp->ev_run[s], hp->s_run));
  // D2H
  SM3_CHECK_CUDA(cudaStreamWaitEvent(hp->s_d2h, hp->ev_run[s], 0));
  SM3_CHECK_CUDA(cudaMemcpyAsync(loss_host, loss, 4, cudaMemcpyDeviceToHost, hp->s_d2h));
  if (dp1_host) {
    SM3_CHECK_CUDA(cudaMemcpyAsync(dp1_host, dp1, hp->in_bytes, cudaMemcpyDeviceToHost, hp->s_d2h));
    SM3_CHECK_CUDA(cudaMemcpyAsync(dp2_host, dp2, hp->in_bytes, cudaMemcpyDeviceToHost, hp->s_d2h));
  }
  SM3_CHECK_CUDA(cudaEventRecord(hp->ev_out[s], hp->s_d2h));
  hp->busy[s] = 1;
  hp->next_ticket = ticket + 1;
  return ticket;
}

extern "C" int sm3_host_pipe_wait(sm3_host_pipe* hp, int64_t ticket) {
  SM3_REQUIRE(hp != nullptr, SM3_ERR_SHAPE, "host_pipe_wait: null handle");
  SM3_REQUIRE(ticket >= 0 && ticket < hp->next_ticket, SM3_ERR_SHAPE, "host_pipe_wait: unknown ticket %lld",
              (long long)ticket);
  // a ticket older than the slot's current occupant was already synchronised by the submit that replaced it
  if (ticket + hp->depth < hp->next_ticket) return SM3_OK;
  SM3_CHECK_CUDA(cudaEventSynchronize(hp->ev_out[ticket % hp->depth]));
  return SM3_OK;
}

// ---------------------------------------------------------------------------------------------------
// peer mode of the pipelined host-buffer entry: the same three-stream pipeline around sm3_infonce_step_peer, so that a
// multi-rank job's host-to-host throughput is not bounded by Python between 0.6 ms steps.  The caller passes the
// symmetric buffers of the slot it wants used and the step's epoch with every submit (skin_sm3_b200.peer owns both).
// ---------------------------------------------------------------------------------------------------
extern "C" size_t sm3_host_pipe_peer_scratch_bytes(int n_local, int n_global, int D, int io_dtype, int depth) {
  if (n_local < 1 || n_global < n_local || !dtype_ok(io_dtype) || depth < 1 || depth > SM3_PIPE_MAX_DEPTH) return 0;
  const size_t step = sm3_infonce_step_peer_scratch_bytes(n_local, n_global, D);
  if (step == 0) return 0;
  const size_t in_bytes = (size_t)n_local * D * dtype_size(io_dtype);
  return align_up(step) + (size_t)depth * pipe_io_stride(in_bytes);
}

extern "C" int sm3_host_pipe_create_peer(sm3_host_pipe** out, int n_local, int n_global, int D, int io_dtype, int depth,
                                         void* device_scratch, size_t scratch_bytes) {
  SM3_REQUIRE(out != nullptr && device_scratch != nullptr, SM3_ERR_SHAPE, "host_pipe_create_peer: null pointer");
  *out = nullptr;
  const size_t need = sm3_host_pipe_peer_scratch_bytes(n_local, n_global, D, io_dtype, depth);
  SM3_REQUIRE(need != 0, SM3_ERR_SHAPE, "host_pipe_create_peer: bad shape / dtype / depth");
  SM3_REQUIRE(scratch_bytes >= need, SM3_ERR_WORKSPACE, "host_pipe_create_peer: scratch %zu < %zu", scratch_bytes, need);
  // build an ordinary handle for a shape whose single-GPU plan is certainly smaller, then re-point it at the peer plan
  sm3_host_pipe* hp = nullptr;
  int rc = sm3_host_pipe_create(&hp, 1, D, io_dtype, SM3_ALGO_AUTO, depth, device_scratch, scratch_bytes);
  if (rc) return rc;
  hp->n_pairs = n_local; hp->n_global = n_global;
  hp->in_bytes = (size_t)n_local * D * dtype_size(io_dtype);
  hp->io_stride = pipe_io_stride(hp->in_bytes);
  hp->step_bytes = align_up(sm3_infonce_step_peer_scratch_bytes(n_local, n_global, D));
  hp->total = need;
  *out = hp;
  return SM3_OK;
}

extern "C" int64_t sm3_host_pipe_submit_peer(sm3_host_pipe* hp, const void* p1_host, const void* p2_host, float temperature,
                                             float* loss_host, void* dp1_host, void* dp2_host, int rank, int world,
                                             void* z_cols_mine, void* const* z_peers_host, void* stats_mine,
                                             void* const* stats_peers_host, void* flags_mine,
                                             void* const* flags_peers_host, unsigned epoch, int mode) {
  SM3_REQUIRE(hp && p1_host && p2_host && loss_host, SM3_ERR_SHAPE, "host_pipe_submit_peer: null pointer");
  SM3_REQUIRE(hp->n_global > hp->n_pairs && hp->n_global == hp->n_pairs * world, SM3_ERR_SHAPE,
              "host_pipe_submit_peer: handle was not created for %d ranks", world);
  SM3_REQUIRE((dp1_host == nullptr) == (dp2_host == nullptr), SM3_ERR_SHAPE,
              "host_pipe_submit_peer: dp1_host/dp2_host must both be given or both NULL");
  SM3_REQUIRE(mode == 0 || (mode >= 2 && mode <= 4), SM3_ERR_SHAPE, "host_pipe_submit_peer: exchange mode must be 0, 2, 3 or 4");
  int dev = -1;
  SM3_CHECK_CUDA(cudaGetDevice(&dev));
  SM3_REQUIRE(dev == hp->device, SM3_ERR_SHAPE, "host_pipe_submit_peer: handle belongs to device %d, current device is %d",
              hp->device, dev);
  const int64_t ticket = hp->next_ticket;
  const int s = (int)(ticket % hp->depth);
  if (hp->busy[s]) SM3_CHECK_CUDA(cudaEventSynchronize(hp->ev_out[s]));
  char* io = hp->base + hp->step_bytes + (size_t)s * hp->io_stride;
  const size_t a = align_up(hp->in_bytes);
  char *p1 = io, *p2 = io + a, *dp1 = io + 2 * a, *dp2 = io + 3 * a;
  float* loss = (float*)(io + 4 * a);
  if (hp->busy[s]) SM3_CHECK_CUDA(cudaStreamWaitEvent(hp->s_h2d, hp->ev_run[s], 0));
  SM3_CHECK_CUDA(cudaMemcpyAsync(p1, p1_host, hp->in_bytes, cudaMemcpyHostToDevice, hp->s_h2d));
  SM3_CHECK_CUDA(cudaMemcpyAsync(p2, p2_host, hp->in_bytes, cudaMemcpyHostToDevice, hp->s_h2d));
  SM3_CHECK_CUDA(cudaEventRecord(hp->ev_in[s], hp->s_h2d));
  SM3_CHECK_CUDA(cudaStreamWaitEvent(hp->s_run, hp->ev_in[s], 0));
  const int rc = sm3_infonce_step_peer(p1, p2, hp->n_pairs, rank, world, hp->D, hp->io_dtype, temperature, 1.0f, loss,
                                       dp1_host ? dp1 : nullptr, dp1_host ? dp2 : nullptr, z_cols_mine, z_peers_host,
                                       stats_mine, stats_peers_host, flags_mine, flags_peers_host, epoch, mode, hp->base,
                                       hp->step_bytes, hp->s_run, hp->s_run);
  if (rc) return rc;
  SM3_CHECK_CUDA(cudaEventRecord(hp->ev_run[s], hp->s_run));
  SM3_CHECK_CUDA(cudaStreamWaitEvent(hp->s_d2h, hp->ev_run[s], 0));
  SM3_CHECK_CUDA(cudaMemcpyAsync(loss_host, loss, 4, cudaMemcpyDeviceToHost, hp->s_d2h));
  if (dp1_host) {
    SM3_CHECK_CUDA(cudaMemcpyAsync(dp1_host, dp1, hp->in_bytes, cudaMemcpyDeviceToHost, hp->s_d2h));
    SM3_CHECK_CUDA(cudaMemcpyAsync(dp2_host, dp2, hp->in_bytes, cudaMemcpyDeviceToHost, hp->s_d2h));
  }
  SM3_CHECK_CUDA(cudaEventRecord(hp->ev_out[s], hp->s_d2h));
  hp->busy[s] = 1;
  hp->next_ticket = ticket + 1;
  return ticket;
}
