#!/usr/bin/env python
"""bench.py -- contrastive pairs/sec (fwd+bwd) of the fused cross-modal InfoNCE hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg4|cfg2|cfg1|cfg3] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" = one pass of the hot path over one synthetic batch: normalise -> fused similarity + InfoNCE
statistics -> CE on the statistics -> fused backward -> normalise-backward (gradients w.r.t. both projector
outputs).  Workloads (BASELINE.json):
  cfg4 (default): global batch N=32768 pairs, D=256, bf16, T=0.1, one cross-modal term -- the configuration the
        north_star target ("global batch 32768 x dim 256") is quoted on; it fits one GPU because the [M,M] logits
        are never materialised.  With N>1 ranks the SAME global batch is row-sharded (strong scaling), negatives
        all-gathered over NCCL.
  cfg2: N=4096 pairs, D=128, bf16 (configs[1]); also always reported inside the cfg4 line under "cfg2".
Rank 0 prints ONE JSON line.  `value` = device-timed whole-job throughput with inputs resident in HBM;
`e2e` = the same metric through the host-buffer C-ABI / public API with H2D + D2H inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

WORKLOADS = {
    "cfg1": dict(n=64, d=128, T=0.1, name="cfg1: SM3 contrastive loss fwd+bwd, batch 64 x dim 128 (the reference's own CPU-runnable case)"),
    "cfg4": dict(n=32768, d=256, T=0.1, name="cfg4: fused cross-modal InfoNCE fwd+bwd, global batch 32768 x dim 256, bf16"),
    "cfg2": dict(n=4096, d=128, T=0.1, name="cfg2: fused cross-modal InfoNCE fwd+bwd, batch 4096 x dim 128, bf16"),
}
SEED = 3407
COMM = os.environ.get("SM3_COMM", "auto")     # multi-rank exchange: auto (peer memory if available) | peer | nccl
# our kernels per step on the production path: normalise (+ NVLink row scatter when sharded), K2, CE on the statistics
# (+ a_j, + statistics scatter), K3, normalise-backward -- one C call -- and the upstream-scale kernel of .backward()
KERNELS_PER_STEP = 6


def ncu_traffic():
    """DRAM bytes per launch of the dominant kernel, from the committed `ncu --set full` capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return dict(tflops=float(p["bf16_tflops"]), tflops_sustained=float(p.get("bf16_tflops_sustained", 0)),
                    hbm=float(p["hbm_gbs"]), source="MEASURED_PEAKS.json (of measured)")
    except Exception:
        return dict(tflops=1590.0, tflops_sustained=1400.0, hbm=6650.0, source="B200_PROFILING.md fallback (of fallback)")


class ClockSampler:
    """Samples SM clock + throttle reasons during the timed region (pynvml; nvidia-smi as a fallback)."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._t = None

    def _loop_nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        names = {
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80): "hw_power_brake_slowdown",
        }
        get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
            mask = get(h)
            for bit, name in names.items():
                if mask & bit:
                    self.reasons.add(name)
            time.sleep(0.002)

    def __enter__(self):
        def run():
            try:
                self._loop_nvml()
            except Exception as e:  # pragma: no cover
                self.reasons.add(f"sampler_error:{type(e).__name__}")
        self._t = threading.Thread(target=run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=2)

    def summary(self):
        s = self.samples
        return {"sm_mhz": (statistics.median(s) if s else None), "sm_min_mhz": (min(s) if s else None),
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def dist_setup(n_gpus: int):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the single JSON line (no NCCL version banner)
        torch.cuda.set_device(local)
        # NCCL prints its version banner on stdout at communicator creation: park fd 1 on stderr until the first
        # collective has run so that stdout carries exactly one line (the JSON).
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    else:
        torch.cuda.set_device(0)
    if n_gpus != world and rank == 0:
        print(f"[bench] note: --gpus {n_gpus} but WORLD_SIZE={world}; using {world} rank(s)", file=sys.stderr)
    return world, rank, local


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x: float, world: int) -> float:
    if world == 1:
        return x
    import torch.distributed as dist
    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def synth(n_global, d, rank, world, device=None, seed=SEED):
    """p1, p2 ~ N(0,1) (the projector ends in an affine-free BatchNorm, simclr.py:26), bf16.  The GLOBAL batch is drawn
    from one seed and every rank takes its slice of pairs, so the N = 1 / 2 / 4 / 8 runs compute the SAME problem and
    their (rank-averaged) losses must agree.  -> (p1_local, p2_local, (P1_global, P2_global))"""
    g = torch.Generator(device="cpu").manual_seed(seed)
    P1 = torch.randn(n_global, d, generator=g).bfloat16()
    P2 = torch.randn(n_global, d, generator=g).bfloat16()
    n_local = n_global // world
    sl = slice(rank * n_local, (rank + 1) * n_local)
    return P1[sl].contiguous(), P2[sl].contiguous(), (P1, P2)


def stage_times(sm3, p1, p2, T, group, world, steps, flush):
    """Per-stage device times of the PRODUCTION path (the one C call `value` times): libsm3's stage-timing facility
    records CUDA events between the kernels inside sm3_infonce_step / sm3_infonce_step_peer.  -> {stage: ms}"""
    import ctypes as C
    lib = sm3.lib()
    a = p1.cuda().requires_grad_(True)
    b = p2.cuda().requires_grad_(True)
    acc, names = {}, []
    lib.sm3_stage_timing(1)
    try:
        for _ in range(max(3, min(steps, 10))):
            if flush is not None:
                flush.add_(1.0)
            a.grad = b.grad = None
            sm3.fused_infonce(a, b, T, precision="bf16", group=group, comm=COMM).backward()
            buf = (C.c_float * 8)()
            k = lib.sm3_stage_timing_read(buf, 8)
            names = [x for x in lib.sm3_stage_timing_names().decode().split(",") if x][:k]
            for nm, v in zip(names, list(buf)[:k]):
                acc.setdefault(nm, []).append(float(v))
    finally:
        lib.sm3_stage_timing(0)
    torch.cuda.synchronize()
    return {k: sum(v) / len(v) for k, v in acc.items()}


def time_device(sm3, p1, p2, T, group, world, steps, warmup, flush, profile=True):
    """K steps of the public op on HBM-resident inputs; per-stage CUDA events on the launching stream."""
    from skin_sm3_b200 import functional as F3
    a = p1.cuda().requires_grad_(True)
    b = p2.cuda().requires_grad_(True)
    marks = []

    def mark(name):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()                      # torch's current stream == the stream the kernels launch on
        marks.append((name, ev))         # `marks` is rebound per step below

    def step():
        a.grad = b.grad = None
        loss = sm3.fused_infonce(a, b, T, precision="bf16", group=group, comm=COMM)
        loss.backward()
        return loss

    for _ in range(warmup):
        step()
    barrier(world)
    # All K steps are enqueued back to back (no host sync inside the timed region, as in a training loop); every
    # step is bracketed by its own CUDA events on the launching stream so the untimed L2 flush stays outside.
    F3._PROFILE = mark if profile else None     # without marks a single-GPU step is ONE C call (sm3_infonce_step)
    recs, loss = [], None
    try:
        for _ in range(steps):
            if flush is not None:
                flush.add_(1.0)          # > L2 (126 MB) write between timed iterations; not timed
            marks = []
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record()
            loss = step()
            e1.record()
            recs.append((e0, e1, marks))
    finally:
        F3._PROFILE = None
    torch.cuda.synchronize()
    per_step, stage_ms = [], {}
    for e0, e1, mk in recs:
        per_step.append(e0.elapsed_time(e1))
        for (n0, ev0), (n1, ev1) in zip(mk[:-1], mk[1:]):
            stage_ms.setdefault(n1, []).append(ev0.elapsed_time(ev1))
    barrier(world)
    total_ms = max_over_ranks(sum(per_step), world)
    stage_avg = {k: sum(v) / len(v) for k, v in stage_ms.items()}
    return total_ms, stage_avg, float(loss.item())


def time_graph(sm3, p1, p2, T, steps, warmup, flush):
    """The same single-GPU step replayed from a CUDA graph (skin_sm3_b200.GraphedInfoNCE): no Python / autograd between
    the kernels.  Device-timed per replay, L2 flushed between replays.  -> ms per step"""
    gr = sm3.GraphedInfoNCE(p1.shape[0], p1.shape[1], T, dtype=torch.bfloat16, precision="bf16")
    gr.p1.copy_(p1.cuda()); gr.p2.copy_(p2.cuda())
    for _ in range(warmup):
        gr.replay()
    recs = []
    for _ in range(steps):
        if flush is not None:
            flush.add_(1.0)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        gr.replay()
        e1.record()
        recs.append((e0, e1))
    torch.cuda.synchronize()
    return sum(a.elapsed_time(b) for a, b in recs) / len(recs)


def time_e2e(sm3, p1, p2, T, group, world, steps, warmup, depth=2):
    """Same metric end to end: every step copies its inputs from pinned host memory (H2D), runs the fused fwd+bwd and
    copies the loss and both gradients back to pinned host memory (D2H), all inside the timed region.  The steps go
    through a `depth`-slot pipeline (copies of neighbouring steps overlap the kernels, as a training loop with a
    prefetching loader does); the one-synchronous-call-per-step figure is returned beside it.
    -> (pipelined ms for `steps`, synchronous ms for `steps`, h2d bytes/step, d2h bytes/step)"""
    n_local, d = p1.shape
    hp1, hp2 = p1.pin_memory(), p2.pin_memory()
    h2d = 2 * n_local * d * 2
    d2h = 4 + h2d

    def timed(run):
        barrier(world)
        t0 = time.perf_counter()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record(); e1.synchronize()
        wall_ms = (time.perf_counter() - t0) * 1e3
        barrier(world)
        return max_over_ranks(max(e0.elapsed_time(e1), wall_ms), world)     # host clock: the copies run on other streams

    pipe = None
    if world == 1:
        host = sm3.HostInfoNCE(n_local, d, torch.bfloat16, sm3.ALGO_AUTO)             # sm3_infonce_host: synchronises
        pipe = sm3.HostInfoNCEPipeline(n_local, d, torch.bfloat16, sm3.ALGO_AUTO, depth)  # sm3_host_pipe_*
    elif COMM != "nccl":
        try:        # peer mode of the same C pipeline: sm3_host_pipe_submit_peer runs the multi-rank fused step
            pipe = sm3.HostInfoNCEPipeline(n_local, d, torch.bfloat16, sm3.ALGO_AUTO, depth, group=group)
        except Exception as e:
            print(f"[bench] peer host pipeline unavailable ({e!r}); e2e falls back to the Python-driven loop", file=sys.stderr)
            pipe = None
    if pipe is not None:
        sink = []

        def run_sync(k=steps):
            for _ in range(k):
                if world == 1:
                    host(hp1, hp2, T)
                else:
                    sink.append(float(pipe.wait(pipe.submit(hp1, hp2, T))[0][0]))

        def run_pipe(k=steps):
            tickets = []
            for i in range(k):
                tickets.append(pipe.submit(hp1, hp2, T))
                if len(tickets) >= depth:
                    sink.append(float(pipe.wait(tickets.pop(0))[0][0]))       # the step's result, read on the host
            for t in tickets:
                sink.append(float(pipe.wait(t)[0][0]))

        run_sync(warmup); run_pipe(warmup)
        ms_sync = timed(run_sync)
        ms_pipe = timed(run_pipe)
        pipe.close()
        return ms_pipe, ms_sync, h2d, d2h

    # multi-rank: the public op on device tensors, fed from / drained to pinned host memory by two copy streams
    main = torch.cuda.current_stream()
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    dev_in = [(torch.empty_like(hp1, device="cuda").requires_grad_(True),
               torch.empty_like(hp2, device="cuda").requires_grad_(True)) for _ in range(depth)]
    host_out = [(torch.empty(1, dtype=torch.float32).pin_memory(), torch.empty_like(hp1).pin_memory(),
                 torch.empty_like(hp2).pin_memory()) for _ in range(depth)]
    ev_run = [None] * depth
    ev_out = [None] * depth
    keep = [None] * depth
    sink = []

    def submit(i):
        s = i % depth
        if ev_out[s] is not None:
            ev_out[s].synchronize()                                     # slot's results have reached the host
            sink.append(float(host_out[s][0][0]))
        a, b = dev_in[s]
        with torch.cuda.stream(s_in), torch.no_grad():
            if ev_run[s] is not None:
                s_in.wait_event(ev_run[s])                              # previous occupant's kernels have read a, b
            a.copy_(hp1, non_blocking=True); b.copy_(hp2, non_blocking=True)
            e_in = torch.cuda.Event(); e_in.record(s_in)
        main.wait_event(e_in)
        a.grad = b.grad = None
        loss = sm3.fused_infonce(a, b, T, precision="bf16", group=group, comm=COMM)
        loss.backward()
        ev_run[s] = torch.cuda.Event(); ev_run[s].record(main)
        keep[s] = (loss, a.grad, b.grad)                                # alive until the D2H below has finished
        with torch.cuda.stream(s_out), torch.no_grad():
            s_out.wait_event(ev_run[s])
            host_out[s][0].copy_(loss.detach().reshape(1), non_blocking=True)
            host_out[s][1].copy_(a.grad, non_blocking=True); host_out[s][2].copy_(b.grad, non_blocking=True)
            ev_out[s] = torch.cuda.Event(); ev_out[s].record(s_out)

    def drain():
        for s in range(depth):
            if ev_out[s] is not None:
                ev_out[s].synchronize()
                ev_out[s] = None
        torch.cuda.synchronize()

    def run_pipe(k=steps):
        for i in range(k):
            submit(i)
        drain()

    def run_sync(k=steps):
        for i in range(k):
            submit(i)
            drain()

    run_pipe(warmup)
    ms_sync = timed(run_sync)
    ms_pipe = timed(run_pipe)
    return ms_pipe, ms_sync, h2d, d2h


def _real_reference():
    """The reference's own code (src/models/simclr.py:290-322 + nn.CrossEntropyLoss, tools/backbone_train.py:531), imported
    unmodified from oracle/_ref/skin_sm3 (vendored by oracle/vendor_ref.py; see there).  None when it is not there."""
    from oracle import vendor_ref
    root = vendor_ref.ref_root()
    if root is None:
        return None
    if root not in sys.path:
        sys.path.insert(0, root)
    try:
        from src.models.simclr import SimCLRSkinV3
    except Exception as e:  # pragma: no cover
        print(f"[bench] reference import failed: {e!r}", file=sys.stderr)
        return None
    import torch.nn as nn
    ident, crit = nn.Identity(), nn.CrossEntropyLoss()

    def step(p1, p2, T):
        a = p1.detach().clone().requires_grad_(True)
        b = p2.detach().clone().requires_grad_(True)
        logits, labels = SimCLRSkinV3._cal_logits(None, a, b, ident, ident, T)     # `self` is unused by the reference
        loss = crit(logits, labels)
        loss.backward()
        return loss.detach(), a.grad, b.grad
    return step


def _ref_sample_pairs(n, d, budget_s):
    """Largest sub-problem (pairs) of the reference's O(M^2)-memory step that fits the host and the time budget: ~7 live
    [M, M] fp32 tensors at the peak, measured 3.0 GB at N = 4096."""
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 16e9
    cap = n
    while cap > 64 and (2 * cap) ** 2 * 4 * 12 > 0.5 * avail:
        cap //= 2
    return cap


def cpu_reference(n, d, T, budget_s=20.0, reps=None, warmup=1):
    """The reference's CPU implementation of the path on the box's host cores, `reps` (+ `warmup`) timed steps of a BOUNDED
    sample: the full step when it is small, else the same step on the largest square sub-problem N_s that fits memory and
    the time budget, scaled to the full configuration by (N / N_s)^2 (the step's work and memory are O(N^2)); the figure is
    then marked "estimated".  kind = "reference" (the real code, vendored) or "port" (oracle/ref_port.py) when absent."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = _real_reference()
    kind = "reference" if step is not None else "port"
    if step is None:
        from oracle import ref_port
        step = ref_port.port_infonce_step
    g = torch.Generator().manual_seed(SEED)
    n_s = _ref_sample_pairs(n, d, budget_s)
    total_reps = (reps or 3) + warmup
    # shrink until (reps + warmup) steps fit the budget: probe a small size, extrapolate quadratically
    probe = min(n_s, 512)
    q1, q2 = torch.randn(probe, d, generator=g), torch.randn(probe, d, generator=g)
    step(q1, q2, T)
    t0 = time.perf_counter(); step(q1, q2, T); t_probe = time.perf_counter() - t0
    while n_s > probe and 1.5 * t_probe * (n_s / probe) ** 2 * total_reps > budget_s:   # 1.5: the step is superquadratic once it leaves the caches
        n_s //= 2
    p1, p2 = torch.randn(n_s, d, generator=g), torch.randn(n_s, d, generator=g)
    for _ in range(warmup):
        step(p1, p2, T)
    times = []
    for _ in range(reps or 3):
        t0 = time.perf_counter()
        step(p1, p2, T)
        times.append(time.perf_counter() - t0)
    t_s = statistics.median(times)
    scale = (n / n_s) ** 2
    what = "SimCLRSkinV3._cal_logits + nn.CrossEntropyLoss, fwd+bwd" if kind == "reference" else "op-for-op port of _cal_logits + CE"
    if n_s == n:
        sample = f"full step N={n} D={d} fp32 ({what}, {cores} threads)"
    else:
        sample = (f"{what} on a square sub-problem N_s={n_s} of N={n} (D={d}, fp32, {cores} threads; the full step needs "
                  f"~{(2 * n) ** 2 * 4 * 7 / 1e9:.0f} GB); time scaled x{scale:.0f} = (N/N_s)^2")
    out = dict(value=n / (t_s * scale), unit="pairs/s", cores=cores, kind=kind, sample=sample,
               sample_pairs=n_s, sample_ms_per_step=t_s * 1e3, reps=len(times), estimated=(n_s != n), scale_factor=scale)
    return out


def heads_probe(sm3, pk):
    """HBM-bound kernels: multi-label head losses (K4 8-head CE, K5 BCE) -- latency at the reference size and
    GB/s at a bandwidth-sized batch -- and K1 row normalisation fwd / bwd."""
    out = {}
    M, D = 1 << 21, 256
    p = torch.randn(M, D, device="cuda", dtype=torch.bfloat16)
    z, inv = sm3.core.normalize_pair(p, None, torch.bfloat16)
    dz = torch.randn(M, D, device="cuda", dtype=torch.float32)
    for op, fn, nbytes in (("l2norm_fwd_2Mx256", lambda: sm3.core.normalize_pair(p, None, torch.bfloat16), M * D * 4 + 4 * M),
                           ("l2norm_bwd_2Mx256", lambda: sm3.core.normalize_bwd(dz, 1, 1.0, z, inv, M, 0, torch.bfloat16),
                            M * D * (4 + 2 + 2) + 4 * M)):
        for _ in range(3):
            fn()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record(); e1.synchronize()
        ms = e0.elapsed_time(e1) / 10
        out[op] = {"us": round(ms * 1e3, 2), "GB/s": round(nbytes / ms / 1e6, 1),
                   "frac_hbm": round(nbytes / ms / 1e6 / pk["hbm"], 3)}
    del p, z, inv, dz
    for name, B in (("b512", 512), ("b4096", 4096), ("b4M", 1 << 22)):     # cfg5 size, cfg2 size, bandwidth size
        x = torch.randn(B, 24, device="cuda", dtype=torch.bfloat16, requires_grad=True)
        y = torch.stack([torch.randint(0, c, (B,), device="cuda") for c in sm3.NUM_CLASSES], 1)
        t = torch.nn.functional.one_hot(y[:, 0], 24).to(torch.bfloat16)
        for op, fn, nbytes in (("ce8", lambda: sm3.multihead_ce(x, y), B * 24 * 4 + B * 64 + 4),
                               ("bce", lambda: sm3.bce_with_logits(x, t), B * 24 * 6)):
            for _ in range(3):
                fn()
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            reps = 20
            e0.record()
            for _ in range(reps):
                fn()
            e1.record(); e1.synchronize()
            ms = e0.elapsed_time(e1) / reps
            out[f"{op}_{name}"] = {"us": round(ms * 1e3, 2), "GB/s": round(nbytes / ms / 1e6, 1),
                                   "frac_hbm": round(nbytes / ms / 1e6 / pk["hbm"], 3)}
    tr = ncu_traffic()
    for k, v in out.items():                      # DRAM bytes per launch from the committed ncu --set full capture
        if k in tr:
            v["algorithmic_bytes"] = tr[k]["algorithmic_bytes"]
            v["traffic"] = tr[k]["dram_bytes_per_launch"]
    return out


def gpu_reference_port(n, d, T, reps=5):
    """Context only: the reference's materialising op sequence (oracle/ref_port.py) on the SAME GPU, fp32."""
    from oracle import ref_port
    g = torch.Generator().manual_seed(SEED)
    p1 = torch.randn(n, d, generator=g).cuda(); p2 = torch.randn(n, d, generator=g).cuda()
    for _ in range(2):
        ref_port.port_infonce_step(p1, p2, T)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        ref_port.port_infonce_step(p1, p2, T)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / reps
    return {"value": n / (ms * 1e-3), "unit": "pairs/s", "ms_per_step": ms,
            "what": f"reference op sequence (port) on this GPU, fp32, N={n} D={d}; not the headline baseline"}


def config_of(wl, n, d, T, world):
    """The `config` object of a bench line: identical for our arm and the reference arm of the same workload / N."""
    return {"workload": wl["name"], "global_pairs": n, "rows_per_rank": 2 * n // world, "dim": d, "temperature": T,
            "terms": 1, "sharding": f"row-block x{world}, global negatives" if world > 1 else "single GPU",
            "l2": "256 MB flush write between timed steps (inputs are L2-sized by design)"}


def comm_description(n, world):
    if world == 1:
        return "none"
    from skin_sm3_b200 import peer as _peer
    if COMM == "nccl":
        return "NCCL all-gather"
    used = "NVLink peer memory (symmetric memory"
    mc = bool(_peer._CACHE) and any(b.multicast for b in _peer._CACHE.values())
    used += ", NVSwitch multicast stores)" if mc else ", unicast stores)"
    if os.environ.get("SM3_PEER_FUSED", "1") != "0" and (n // world) % 128 == 0:
        used += ", fused exchange: scatter+signal in the producer kernels, waits inside K2/K3"
        from skin_sm3_b200.functional import _fused_mode
        mode = _fused_mode()
        used += {3: "; rows pushed from inside K2 (mode 3)",
                 4: "; symmetric forward across ranks: W/2 of W column blocks per rank, column sums shipped to the row "
                    "owners (mode 4)"}.get(mode, " (mode 2)")
    return used


def parity_block(sm3, p1, p2, full, n, d, T, group, world, rank):
    """Checked OUTSIDE every timed region, on the production path and at the benchmarked size: the loss (rank mean) and the
    gradient of a row sample of rank 0's shard against oracle.infonce_rowblock (fp32-GEMM mode: pinned to the reference
    goldens to 1e-5, tests/test_oracle_golden.py), and, when sharded, the fused NVLink exchange against the NCCL path.
    The oracle is the checker here, never the thing measured."""
    import numpy as np
    a = p1.cuda().requires_grad_(True)
    b = p2.cuda().requires_grad_(True)
    loss = sm3.fused_infonce(a, b, T, precision="bf16", group=group, comm=COMM)
    loss.backward()
    out = {"loss_local": float(loss)}
    gl = loss.detach().clone().double()
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(gl)
        gl /= world
        a2 = p1.cuda().requires_grad_(True)
        b2 = p2.cuda().requires_grad_(True)
        l2 = sm3.fused_infonce(a2, b2, T, precision="bf16", group=group, comm="nccl")
        l2.backward()
        e = max(float((a.grad.float() - a2.grad.float()).abs().max() / a2.grad.float().abs().max()),
                float((b.grad.float() - b2.grad.float()).abs().max() / b2.grad.float().abs().max()))
        out["fused_vs_nccl_grad_relerr"] = max_over_ranks(e, world)
        out["fused_vs_nccl_loss_relerr"] = max_over_ranks(abs(float(loss) - float(l2)) / abs(float(l2)), world)
    out["loss"] = float(gl)
    if rank == 0:
        from oracle import sm3_oracle as O                      # checker only
        nl = n // world
        rng = np.random.default_rng(11)
        loc = np.unique(np.concatenate([np.arange(32), np.arange(nl - 32, nl), rng.choice(nl, min(nl, 64), replace=False)]))
        rows = np.concatenate([loc, n + loc])                   # rank 0 owns pairs [0, nl): global rows loc and n + loc
        t0 = time.perf_counter()
        ref_loss, ref_dp, _ = O.infonce_rowblock(full[0].float().numpy(), full[1].float().numpy(), T, rows,
                                                 matmul_dtype=np.float32, threads=os.cpu_count())
        got = torch.cat([a.grad[torch.from_numpy(loc).cuda()], b.grad[torch.from_numpy(loc).cuda()]]).float().cpu().numpy()
        # the sharded op returns d(sum over ranks of the per-rank mean losses)/dp = world x d(global mean loss)/dp
        got = got / world
        out.update({"loss_oracle_rowblock": ref_loss, "loss_relerr": abs(float(gl) - ref_loss) / abs(ref_loss),
                    "grad_relerr_rowblock": float(np.abs(got - ref_dp).max() / np.abs(ref_dp).max()),
                    "rows_checked": int(len(rows)), "tolerance": 2e-2, "oracle_seconds": round(time.perf_counter() - t0, 1),
                    "oracle": "oracle.sm3_oracle.infonce_rowblock (fp32 GEMM, fp64 accumulation) on the same bf16-rounded inputs"})
        out["ok"] = bool(out["loss_relerr"] <= 2e-2 and out["grad_relerr_rowblock"] <= 2e-2 and
                         out.get("fused_vs_nccl_grad_relerr", 0.0) <= 1e-2)
    barrier(world)
    return out


def run_ours(args):
    import skin_sm3_b200 as sm3
    world, rank, local = dist_setup(args.gpus)
    group = None
    if world > 1:
        import torch.distributed as dist
        group = dist.group.WORLD
    wl = WORKLOADS[args.workload]
    n, d, T = wl["n"], wl["d"], wl["T"]
    assert n % world == 0
    pk = peaks()
    p1, p2, full = synth(n, d, rank, world)
    flush = torch.zeros(64 * 1024 * 1024, device="cuda")            # 256 MB > 126 MB L2
    with ClockSampler(local) as clk:
        # `value`: the production path (the whole step enqueued by ONE C call, no event marks, no Python between kernels)
        total_ms, _, _ = time_device(sm3, p1, p2, T, group, world, args.steps, args.warmup, flush, profile=False)
    e2e_ms, e2e_sync_ms, h2d, d2h = time_e2e(sm3, p1, p2, T, group, world, args.steps, args.warmup)
    # per-kernel durations of the SAME production path (events recorded inside the C call), a separate pass
    stages = stage_times(sm3, p1, p2, T, group, world, args.steps, flush)
    stage_source = "events inside sm3_infonce_step%s (production path)" % ("" if world == 1 else "_peer")
    if not stages:       # NCCL / unfused exchange: no single C call -> Python-composed sequence of the same kernels
        _, stages, _ = time_device(sm3, p1, p2, T, group, world, max(3, args.steps // 2), 2, flush, profile=True)
        stages = {{"stats_fwd": "infonce_fwd", "stats_bwd": "infonce_bwd"}.get(k, k): v for k, v in stages.items()}
        stage_source = "CUDA events between the Python-composed kernels (exchange path without a one-call step)"
    parity = parity_block(sm3, p1, p2, full, n, d, T, group, world, rank)
    ms_step = total_ms / args.steps
    m_cols, m_rows = 2 * n, 2 * n // world
    comm_used = comm_description(n, world)
    flops_bwd = 4.0 * m_rows * m_cols * d
    flops_fwd = 2.0 * m_rows * m_cols * d
    line = {
        "metric": "contrastive pairs/sec (fwd+bwd)", "value": n / (ms_step * 1e-3), "unit": "pairs/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": config_of(wl, n, d, T, world),
        "exchange": comm_used,
        "loss": parity["loss"],
        "e2e": {"value": n / (e2e_ms / args.steps * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps,
                "sync_value": n / (e2e_sync_ms / args.steps * 1e-3), "sync_ms_per_step": e2e_sync_ms / args.steps,
                "api": ("sm3_host_pipe_submit/wait (C ABI, pinned host buffers, 2-slot copy/compute pipeline); "
                        "sync_value = one synchronous sm3_infonce_host call per step") if world == 1 else E2E_MULTI_API},
        "gpu_launches": (KERNELS_PER_STEP + (1 if world > 1 and "mode 4" in comm_used else 0)) * args.steps,
        "stages_ms": {k: round(v, 4) for k, v in stages.items()},
        "clocks": clk.summary(),
        "parity": parity,
    }
    tr = ncu_traffic()
    t_bwd = stages.get("infonce_bwd")
    t_fwd = stages.get("infonce_fwd") or stages.get("infonce_fwd_sym") or stages.get("infonce_fwd_push")
    if t_bwd:
        ach = flops_bwd / (t_bwd * 1e-3) / 1e12
        kname = "infonce_tc_bwd_kernel" if d > 128 else "infonce_tc_bwd2_kernel"
        line["roofline"] = {"bound": "tensor", "kernel": kname + " (K3)" + (", incl. in-kernel waits for the peers' statistics" if world > 1 else ""),
                            "achieved": ach, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": ach / pk["tflops"],
                            "traffic": (tr.get(args.workload, {}).get("bwd_dram_bytes_per_launch") if world == 1 else None),
                            "traffic_source": tr.get("source"), "stage_source": stage_source,
                            "peak_source": pk["source"] + ", burst bf16", "flops_per_launch": flops_bwd,
                            "launch_ms": t_bwd, "frac_of_sustained": ach / pk["tflops_sustained"] if pk["tflops_sustained"] else None}
        r = line["roofline"]
        if t_fwd:
            achf = flops_fwd / (t_fwd * 1e-3) / 1e12
            sym, ex_fl = fwd_symmetric(sm3, n, d) if world == 1 else (False, flops_fwd)
            r["fwd"] = {"kernel": ("infonce_tc_fwdsym_kernel (K2, upper-triangular tiles)" if sym else
                                   "infonce_tc_fwdsym_mr_kernel (K2, W/2 of the W column blocks)" if "mode 4" in comm_used else
                                   "infonce_tc_fwd2_kernel (K2)") +
                                  (", incl. in-kernel waits for the peers' rows" if world > 1 else ""),
                        "achieved": achf, "frac": achf / pk["tflops"], "launch_ms": t_fwd, "flops_per_launch": flops_fwd,
                        "executed_flops_per_launch": ex_fl,
                        "note": ("algorithmic 2 M^2 D as SURVEY 8d defines it; the symmetric kernel EXECUTES about half of it, "
                                 "so frac can exceed what the tensor pipe did") if sym else None,
                        "traffic": (tr.get(args.workload, {}).get("fwd_dram_bytes_per_launch") if world == 1 else None)}
        r["step"] = {"achieved": (flops_fwd + flops_bwd) / (ms_step * 1e-3) / 1e12,
                     "frac": (flops_fwd + flops_bwd) / (ms_step * 1e-3) / 1e12 / pk["tflops"],
                     "what": "algorithmic 6 M^2 D / W per rank over the whole timed step"}
        r["stages_ms"] = line["stages_ms"]
        r["parity"] = {k: parity.get(k) for k in ("ok", "loss", "loss_oracle_rowblock", "loss_relerr", "grad_relerr_rowblock",
                                                  "fused_vs_nccl_grad_relerr", "rows_checked", "tolerance") if k in parity}
    line["step_tc_frac"] = (flops_fwd + flops_bwd) / (ms_step * 1e-3) / 1e12 / pk["tflops"]
    if world == 1:
        try:
            gms = time_graph(sm3, p1, p2, T, args.steps, args.warmup, flush)
            line["cuda_graph"] = {"ms_per_step": gms, "value": n / (gms * 1e-3), "unit": "pairs/s",
                                  "what": "same step replayed from a CUDA graph (GraphedInfoNCE), device-timed"}
        except Exception as e:
            line["cuda_graph"] = {"error": repr(e)[:300]}
    if rank == 0 and world == 1 and not args.no_extras:
        line["cpu_baseline"] = cpu_reference(n, d, T, budget_s=25.0, reps=2)
        r = line.get("roofline", {})
        if args.workload != "cfg2":     # configs[1] rides along, inside `roofline` so the driver's record keeps it
            try:
                r["cfg2"] = cfg2_block(sm3, pk, args, flush)
            except Exception as e:
                r["cfg2"] = {"error": repr(e)[:300]}
        try:
            line["cfg4_sweep"] = cfg4_sweep(sm3, args.steps, flush)
            r["cfg4_sweep"] = line["cfg4_sweep"]
        except Exception as e:
            line["cfg4_sweep"] = {"error": repr(e)[:300]}
        try:
            r["hbm_kernels"] = heads_probe(sm3, pk)
        except Exception as e:   # the head probe must never take the headline down
            r["hbm_kernels"] = {"error": repr(e)}
        # The remaining probes run in a CHILD process with a timeout: whatever happens there (exception, device trap, hang)
        # cannot touch the headline numbers above or keep this process from printing its line.
        import subprocess
        try:
            rr = subprocess.run([sys.executable, os.path.abspath(__file__), "--extras-child"], capture_output=True,
                                text=True, timeout=150)
            got = _last_json(rr.stdout)
            r.update(got if got is not None else {"extras_child": {"error": (rr.stderr or "no output")[-300:]}})
        except subprocess.TimeoutExpired as e:
            got = _last_json(e.stdout.decode() if isinstance(e.stdout, bytes) else e.stdout)
            r.update(got or {})
            r["extras_child"] = {"error": "timeout after 150 s"}
        except Exception as e:
            r["extras_child"] = {"error": repr(e)[:300]}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


E2E_MULTI_API = ("sm3_host_pipe_submit_peer/wait (C ABI, pinned host buffers, 2-slot copy/compute pipeline around "
                 "sm3_infonce_step_peer); sync_value = waited for after every submit.  (SM3_COMM=nccl: the public op fed "
                 "from / drained to pinned host memory by two copy streams in Python)")


def fwd_symmetric(sm3, n, d):
    """(does the single-rank forward of this shape run the symmetric kernel, FLOPs it executes per launch)."""
    import ctypes as C
    from skin_sm3_b200 import _lib
    tiles = C.c_longlong(0)
    on = bool(_lib.lib().sm3_infonce_fwd_symmetric(int(n), int(d), C.byref(tiles)))
    return on, float(tiles.value) * 2.0 * 128 * 128 * d


def cfg2_block(sm3, pk, args, flush):
    """BASELINE configs[1] (4096 x 128): production step (eager op), CUDA-graph replay, e2e, per-kernel rooflines from the
    production path's stage events, full-size parity against the fp64 closed form, the reference on the host cores and
    the reference's op sequence on this GPU."""
    import numpy as np
    w2 = WORKLOADS["cfg2"]
    n2, d2, T2 = w2["n"], w2["d"], w2["T"]
    q1, q2, _ = synth(n2, d2, 0, 1)
    ms2, _, _ = time_device(sm3, q1, q2, T2, None, 1, args.steps, args.warmup, flush, profile=False)
    st2 = stage_times(sm3, q1, q2, T2, None, 1, args.steps, flush)
    e2, e2s, _, _ = time_e2e(sm3, q1, q2, T2, None, 1, args.steps, args.warmup)
    try:
        g2 = time_graph(sm3, q1, q2, T2, args.steps, args.warmup, flush)
    except Exception as e:
        g2 = None
        print(f"[bench] cfg2 graph replay failed: {e!r}", file=sys.stderr)
    m2 = 2 * n2
    f2 = 6.0 * m2 ** 2 * d2
    tr = ncu_traffic().get("cfg2", {})
    blk = {"workload": w2["name"], "value": n2 / (ms2 / args.steps * 1e-3), "unit": "pairs/s",
           "ms_per_step": ms2 / args.steps, "e2e_value": n2 / (e2 / args.steps * 1e-3),
           "e2e_sync_value": n2 / (e2s / args.steps * 1e-3),
           "cuda_graph_value": (n2 / (g2 * 1e-3)) if g2 else None, "cuda_graph_ms_per_step": g2,
           "step_frac": f2 / (ms2 / args.steps * 1e-3) / 1e12 / pk["tflops"],
           "cuda_graph_step_frac": (f2 / (g2 * 1e-3) / 1e12 / pk["tflops"]) if g2 else None,
           "stages_ms": {k: round(v, 4) for k, v in st2.items()}}
    for key, nm, fl in (("bwd", "infonce_bwd", 4.0 * m2 * m2 * d2), ("fwd", "infonce_fwd", 2.0 * m2 * m2 * d2)):
        if st2.get(nm):
            ach = fl / (st2[nm] * 1e-3) / 1e12
            blk[key] = {"launch_us": round(st2[nm] * 1e3, 2), "achieved": ach, "frac": ach / pk["tflops"], "unit": "TFLOP/s",
                        "flops_per_launch": fl, "traffic": tr.get(key + "_dram_bytes_per_launch")}
            if key == "fwd":
                blk[key]["symmetric"], blk[key]["executed_flops_per_launch"] = fwd_symmetric(sm3, n2, d2)
    # parity at this size: every gradient element against the fp64 closed form (the checker, outside all timing)
    from oracle import sm3_oracle as O
    a, b = q1.cuda().requires_grad_(True), q2.cuda().requires_grad_(True)
    loss = sm3.fused_infonce(a, b, T2, precision="bf16")
    loss.backward()
    ref_loss, r1, r2 = O.infonce_closed_form(q1.float().numpy(), q2.float().numpy(), T2, chunk=2048)
    blk["parity"] = {"loss": float(loss), "loss_oracle": ref_loss, "loss_relerr": abs(float(loss) - ref_loss) / abs(ref_loss),
                     "grad_relerr_all_rows": float(max(np.abs(a.grad.float().cpu().numpy() - r1).max() / np.abs(r1).max(),
                                                       np.abs(b.grad.float().cpu().numpy() - r2).max() / np.abs(r2).max())),
                     "tolerance": 2e-2}
    blk["cpu_baseline"] = cpu_reference(n2, d2, T2, budget_s=12.0, reps=2)
    try:
        blk["gpu_reference_port"] = gpu_reference_port(n2, d2, T2)
    except Exception as e:
        blk["gpu_reference_port"] = {"error": repr(e)[:200]}
    return blk


def cfg4_sweep(sm3, steps, flush, d=256, T=0.1):
    """SURVEY 8d config 4 on one GPU: N in {8192, 16384, 32768}, D = 256, a single term and the four style-0 terms
    (derm + clinic + 0.5 cross + 0.5 cross) enqueued by one grouped call.  Device-timed, inputs resident."""
    out = []
    for n in (8192, 16384, 32768):
        g = torch.Generator().manual_seed(SEED + n)
        pairs = [(torch.randn(n, d, generator=g).bfloat16().cuda().requires_grad_(True),
                  torch.randn(n, d, generator=g).bfloat16().cuda().requires_grad_(True)) for _ in range(4)]
        row = {"global_pairs": n, "dim": d}
        for name, fn, terms in (("one_term", lambda: sm3.fused_infonce(pairs[0][0], pairs[0][1], T, precision="bf16"), 1),
                                ("four_terms", lambda: sm3.fused_infonce_multi(pairs, T, [1, 1, 0.5, 0.5], precision="bf16"), 4)):
            for _ in range(2):
                fn().backward()
            ms = []
            for _ in range(max(3, min(steps, 8))):
                for a, b in pairs:
                    a.grad = b.grad = None
                flush.add_(1.0)
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                fn().backward()
                e1.record()
                ms.append((e0, e1))
            torch.cuda.synchronize()
            t = sum(a.elapsed_time(b) for a, b in ms) / len(ms)
            flops = terms * 6.0 * (2 * n) ** 2 * d
            row[name] = {"ms_per_step": round(t, 4), "pairs_per_s": round(n / (t * 1e-3), 1),
                         "tc_frac_of_burst_peak": round(flops / (t * 1e-3) / 1e12 / peaks()["tflops"], 3)}
        out.append(row)
        del pairs
    return out


def _last_json(text):
    """Last line of `text` that parses as a JSON object (a crashed child may leave a truncated final line)."""
    for l in reversed((text or "").splitlines()):
        if l.startswith("{"):
            try:
                return json.loads(l)
            except ValueError:
                continue
    return None


def _ev_us(fn, reps=30, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def small_shapes_probe(sm3):
    """The reference's REAL batch sizes (run.sh: 96 pairs over 2 GPUs = 48 per rank; 256 per rank in cfg3), D = 128: the
    step is launch-bound, so what matters is launches and host time.  Per size: the eager op (Python + autograd + one C
    call), the CUDA-graph replay, the four style-0 terms in one grouped call, and the reference's materialising op
    sequence (oracle/ref_port.py, context only) on the same GPU.  Microseconds per step, back-to-back launches."""
    from oracle import ref_port
    out = []
    for n in (48, 256, 1024):
        d, T = 128, 0.1
        g = torch.Generator().manual_seed(SEED + n)
        pairs = [(torch.randn(n, d, generator=g).bfloat16().cuda().requires_grad_(True),
                  torch.randn(n, d, generator=g).bfloat16().cuda().requires_grad_(True)) for _ in range(4)]
        a, b = pairs[0]
        gr = sm3.GraphedInfoNCE(n, d, T, dtype=torch.bfloat16, precision="bf16")
        gr.p1.copy_(a.detach()); gr.p2.copy_(b.detach())
        a32, b32 = a.detach().float(), b.detach().float()

        def eager():
            a.grad = b.grad = None
            sm3.fused_infonce(a, b, T, precision="bf16").backward()

        def grouped():
            sm3.fused_infonce_multi(pairs, T, [1, 1, 0.5, 0.5], precision="bf16").backward()

        out.append({"pairs": n, "dim": d,
                    "eager_us": round(_ev_us(eager), 1), "cuda_graph_us": round(_ev_us(gr.replay), 1),
                    "four_terms_grouped_us": round(_ev_us(grouped), 1),
                    "reference_port_gpu_us": round(_ev_us(lambda: ref_port.port_infonce_step(a32, b32, T), reps=10), 1)})
    return out


def tc_kernel_probe(sm3):
    """K2 / K3 launch times at cfg2 (4096 x 128: the D = 128 regime of every reference config) and at N = 8192 x 256 from the
    production path's stage events, for the kernel variants selectable at run time: backward form 1 (softmax warps split
    the tile's columns) vs 2 (tile-alternating groups, a_j through shared memory) vs 3 (2 with 128-column tiles, D <= 128),
    FMA-pipe exponentials in either kernel.  Microseconds per launch."""
    knobs = ("SM3_TC_FWD_BM", "SM3_TC_POLY", "SM3_TC_BWD_NS", "SM3_TC_BWD_V", "SM3_TC_BWD_POLY")
    saved = {k: os.environ.get(k) for k in knobs}
    out = {}
    try:
        for n, d in ((4096, 128), (8192, 256)):
            p1, p2, _ = synth(n, d, 0, 1, seed=SEED + n)
            for name, cfg in (("default", {}), ("bwd_v1", {"SM3_TC_BWD_V": "1"}), ("bwd_v2", {"SM3_TC_BWD_V": "2"}),
                              ("bwd_v3_poly0", {"SM3_TC_BWD_V": "3", "SM3_TC_BWD_POLY": "0"}),
                              ("bwd_v3_poly2", {"SM3_TC_BWD_V": "3", "SM3_TC_BWD_POLY": "2"}),
                              ("bwd_v4_poly0", {"SM3_TC_BWD_V": "4", "SM3_TC_BWD_POLY": "0"}),
                              ("bwd_v4_poly2", {"SM3_TC_BWD_V": "4", "SM3_TC_BWD_POLY": "2"}),
                              ("fwd_poly0", {"SM3_TC_POLY": "0"}), ("fwd_poly2", {"SM3_TC_POLY": "2"})):
                for k in knobs:
                    os.environ.pop(k, None)
                os.environ.update(cfg)
                sm3.reload_env()
                st = stage_times(sm3, p1, p2, 0.1, None, 1, 10, None)
                out[f"n{n}_d{d}_{name}"] = {"fwd_us": round(st.get("infonce_fwd", 0) * 1e3, 1),
                                            "bwd_us": round(st.get("infonce_bwd", 0) * 1e3, 1)}
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        sm3.reload_env()
    return out


def kmeans_probe(sm3):
    """N4: one DeepCluster clustering (10 iterations + final assignment, K = 5, D = 512) at the Derm7pt bank size and at
    100k samples; milliseconds, host-timed with a device sync on both sides (the function never syncs itself)."""
    out = {}
    for n in (413, 100000):
        emb = torch.nn.functional.normalize(torch.randn(n, 512, device="cuda"), dim=1)
        init = torch.randperm(n)[:5].cuda()
        for _ in range(2):
            sm3.spherical_kmeans(emb, init)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            sm3.spherical_kmeans(emb, init)
        torch.cuda.synchronize()
        out[f"n{n}_ms"] = round((time.perf_counter() - t0) / 5 * 1e3, 3)
    return out


def head_block_probe(sm3):
    """N3 tail at the run.sh shape (B = 256 per GPU, 8 feature slots x 512, 24 prototype rows, DeepCluster CE with
    ignore_index): fused normalise + prototype heads + 8-head CE, forward + backward, against the reference's op sequence
    (tools/mlc_train.py:81-87 and :255-261 -- 8 normalise + 8 Linear + 8 CE and their backward) on the same GPU.
    Microseconds per step, back-to-back launches; plus a bandwidth-sized batch against the HBM roofline."""
    import torch.nn.functional as F
    counts = [5, 3, 2, 3, 3, 3, 3, 2]
    out = {}
    for b in (256, 65536):
        g = torch.Generator().manual_seed(SEED + b)
        sa = torch.randn(8, b, 512, generator=g).cuda().requires_grad_(True)
        ws = [torch.randn(n, 512, generator=g).mul_(0.05).cuda().requires_grad_(True) for n in counts]
        tg = torch.stack([torch.randint(0, n, (b,), generator=g) for n in counts], dim=1).cuda()
        tg[::7, 3] = -100
        crit = torch.nn.CrossEntropyLoss(ignore_index=-100)

        def ours():
            _, lg = sm3.proto_heads(sa, ws, True)
            sm3.multihead_ce(lg, tg, temperature=0.5, ignore_index=-100, class_counts=counts).backward()

        def ref():
            z = [F.normalize(sa[i], dim=-1, p=2) for i in range(8)]
            loss = 0
            for i in range(8):
                loss = loss + crit((z[i] @ ws[i].t()) / 0.5, tg[:, i])
            (loss / 8).backward()

        t_o, t_r = _ev_us(ours, reps=20), _ev_us(ref, reps=20)
        key = f"b{b}"
        out[key] = {"fused_us": round(t_o, 1), "reference_ops_us": round(t_r, 1)}
        if b > 10000:      # forward reads + writes the features, backward reads z and writes d feats
            byts = 4.0 * 8 * b * 512 * 4
            out[key]["fused_GBps"] = round(byts / (t_o * 1e-6) / 1e9, 1)
    return out


def run_cfg3(args):
    """the pretraining step of tools/backbone_train.py:95-130 (style 0) on synthetic 224 x 224 pairs,
    256 pairs per GPU, dual ResNet-50 branches (stock torchvision / cuDNN, out of scope) + the drop-in SimCLRSkinV32 whose
    four loss terms run on the fused kernels; with N > 1 ranks: SyncBatchNorm + DDP as the script does
    (backbone_train.py:510,522) and global negatives.  Timed twice: fused loss path, and the reference's materialising
    op sequence (oracle/ref_port.py on the GPU) in the same model -- the loss is a small share of this step."""
    import torch.nn as nn
    world, rank, local = dist_setup(args.gpus)
    from skin_sm3_b200 import dropin, functional as F3
    dropin.install()
    from src.models.simclr import SimCLRSkinV32
    n_local, T, d = 256, 0.1, 128
    torch.manual_seed(SEED)
    model = SimCLRSkinV32("resnet50", weights=None, proj_dim=d, temperature=T).cuda().to(memory_format=torch.channels_last)
    if world > 1:
        os.environ["SM3_GLOBAL_NEGATIVES"] = "1"
        model = nn.SyncBatchNorm.convert_sync_batchnorm(model)
        model = nn.parallel.DistributedDataParallel(model, device_ids=[local])
    opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9, weight_decay=1e-6)
    crit = nn.CrossEntropyLoss().cuda()
    g = torch.Generator().manual_seed(SEED + rank)
    imgs = [torch.randn(n_local, 3, 224, 224, generator=g).cuda().to(memory_format=torch.channels_last) for _ in range(4)]

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            outs = model([imgs[0], imgs[1]], [imgs[2], imgs[3]], 0)
            loss = crit(*outs[0]) + crit(*outs[1]) + 0.5 * crit(*outs[2][0]) + 0.5 * crit(*outs[2][1])
        loss.backward()
        opt.step()
        return loss

    def timed(k):
        for _ in range(2):
            step()
        barrier(world)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            loss = step()
        e1.record(); e1.synchronize()
        barrier(world)
        return max_over_ranks(e0.elapsed_time(e1), world) / k, float(loss)

    k = max(2, min(args.steps, 6))
    with ClockSampler(local) as clk:
        ms_fused, loss_fused = timed(k)
    line = {"metric": "contrastive pairs/sec (fwd+bwd)", "value": n_local * world / (ms_fused * 1e-3), "unit": "pairs/s",
            "n_gpus": world, "steps": k, "warmup": 2, "ms_per_step": ms_fused, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16 autocast", "data": "synthetic",
            "config": {"workload": "cfg3: SM3 pretraining step, dual ResNet-50 branches, 224x224, 256 pairs/GPU, style 0 "
                                   "(4 InfoNCE terms), SGD step included", "pairs_per_gpu": n_local, "dim": d,
                       "temperature": T, "backbone": "torchvision resnet50 (stock cuDNN, out of scope)"},
            "loss": loss_fused, "clocks": clk.summary()}
    if world == 1 and not args.no_extras:
        from oracle import ref_port                      # context only: the reference's op sequence in the same model
        orig = F3.cal_logits
        F3.cal_logits = lambda p1, p2, temperature, precision="auto", group=None: ref_port.port_cal_logits(p1.float(), p2.float(), temperature)
        try:
            ms_ref, loss_ref = timed(k)
        finally:
            F3.cal_logits = orig
        line["reference_loss_path_same_model"] = {"ms_per_step": ms_ref, "loss": loss_ref,
                                                  "loss_path_saving_ms": ms_ref - ms_fused,
                                                  "what": "identical step with the reference's materialising _cal_logits op "
                                                          "sequence (port, fp32) in place of the fused kernels"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path -- the real SimCLRSkinV3._cal_logits +
    nn.CrossEntropyLoss, unmodified, from oracle/_ref (oracle/ref_port.py only when that is absent) -- on all host threads,
    exactly --steps timed steps after --warmup, each a bounded sample of our arm's workload (see cpu_reference)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    wl = WORKLOADS["cfg4" if args.workload == "cfg3" else args.workload]
    n, d, T = wl["n"], wl["d"], wl["T"]
    steps, warmup = max(1, args.steps), max(0, args.warmup)
    r = cpu_reference(n, d, T, budget_s=150.0, reps=steps, warmup=warmup)
    line = {
        "impl": "reference", "metric": "contrastive pairs/sec (fwd+bwd)", "value": r["value"], "unit": "pairs/s",
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": r["sample_ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_of(wl, n, d, T, world),
        "estimated": r["estimated"], "scale_factor": r["scale_factor"],
        "note": ("value = pairs/s of the full configuration; ms_per_step = measured time of one sampled step"
                 + (" (value is an ESTIMATE: sample time x scale_factor)" if r["estimated"] else "")),
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "estimated", "scale_factor")},
        "e2e": {"value": r["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "device": "host CPU",
    }
    print(json.dumps(line))


def retrieval_probe(sm3):
    """N1: fused similarity + top-k (sm3_sim_topk) against `q @ bank.T; topk` on the same GPU, at the KNN evaluator's
    shape (src/models/evaluator.py:43-83: k = 200) and at a cfg2-sized in-batch retrieval."""
    out = {}
    for name, bq, nb, d, k in (("knn_eval_b512_bank16k_k200", 512, 16384, 128, 200), ("in_batch_8192_k5", 8192, 8192, 128, 5)):
        q = torch.nn.functional.normalize(torch.randn(bq, d, device="cuda"), dim=1)
        bank = torch.nn.functional.normalize(torch.randn(nb, d, device="cuda"), dim=1)
        ours = _ev_us(lambda: sm3.sim_topk(q, bank, k), reps=10)
        ref = _ev_us(lambda: (q @ bank.T).topk(k, dim=-1), reps=10)
        v, i = sm3.sim_topk(q, bank, k)
        rv, ri = (q @ bank.T).topk(k, dim=-1)
        out[name] = {"ours_us": round(ours, 1), "matmul_topk_us": round(ref, 1),
                     "index_agreement": float((i == ri).float().mean())}
        saved = os.environ.get("SM3_TOPK_TILED")
        try:        # the three forms side by side: 0 single pass, 1 tiled + threshold filter, 2 materialised + radix select
            for form in ("0", "1", "2"):
                os.environ["SM3_TOPK_TILED"] = form
                out[name][f"form{form}_us"] = round(_ev_us(lambda: sm3.sim_topk(q, bank, k), reps=5), 1)
        finally:
            if saved is None:
                os.environ.pop("SM3_TOPK_TILED", None)
            else:
                os.environ["SM3_TOPK_TILED"] = saved
    return out


def projector_tail_probe(sm3):
    """N2: the last Linear(2048 -> D, no bias) + BatchNorm1d(D, affine=False) + F.normalize of make_projector
    (src/models/simclr.py:25-26, :294) at the cfg2 row count, bf16: stock torch layers + K1 against the fused tail
    (tcgen05 GEMM with the statistics in its epilogue, then BatchNorm + L2 in one pass), forward and forward + backward."""
    import ctypes as C
    pk = peaks()
    out = {}
    for r, k, d in ((8192, 2048, 128), (65536, 2048, 256)):
        h = torch.relu(torch.randn(r, k, device="cuda")).bfloat16().requires_grad_(True)
        lin = torch.nn.Linear(k, d, bias=False).cuda()
        bn = torch.nn.BatchNorm1d(d, affine=False).cuda()
        w16 = lin.weight.detach().bfloat16().requires_grad_(True)
        gz = torch.randn(r, d, device="cuda")

        def stock_fwd():
            return sm3.l2_normalize(bn(torch.nn.functional.linear(h, w16)).float(), out_dtype=torch.bfloat16)

        lib = sm3.lib()
        y = torch.empty((r, d), dtype=torch.float32, device="cuda")
        totals = torch.empty(2 * d, dtype=torch.float32, device="cuda")
        ws = torch.empty(int(lib.sm3_proj_tail_workspace_bytes(r, d)), dtype=torch.uint8, device="cuda")
        mean, rstd = torch.empty(d, device="cuda"), torch.empty(d, device="cuda")
        z = torch.empty((r, d), dtype=torch.bfloat16, device="cuda")
        inv = torch.empty(r, device="cuda")
        st = torch.cuda.current_stream().cuda_stream

        def gemm():
            lib.sm3_proj_tail_gemm(h.data_ptr(), w16.data_ptr(), r, k, d, 2, y.data_ptr(), totals.data_ptr(), ws.data_ptr(),
                                   ws.numel(), st)

        def fused_fwd():
            gemm()
            lib.sm3_proj_tail_bn_l2(y.data_ptr(), r, d, totals.data_ptr(), float(r), 1e-5, 1e-12, 1, 0.1,
                                    bn.running_mean.data_ptr(), bn.running_var.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                                    z.data_ptr(), inv.data_ptr(), st)

        t_gemm = _ev_us(gemm, reps=20)
        t_cublas = _ev_us(lambda: torch.nn.functional.linear(h, w16), reps=20)
        nbytes = r * k * 2 + d * k * 2 + r * d * 4
        out[f"r{r}_k{k}_d{d}"] = {
            "stock_fwd_us": round(_ev_us(stock_fwd, reps=20), 1), "fused_fwd_us": round(_ev_us(fused_fwd, reps=20), 1),
            "tail_gemm_us": round(t_gemm, 1), "cublas_linear_us": round(t_cublas, 1),
            "tail_gemm_GBps": round(nbytes / t_gemm / 1e3, 1), "tail_gemm_frac_hbm": round(nbytes / t_gemm / 1e3 / pk["hbm"], 3),
            "tail_gemm_TFLOPs": round(2.0 * r * k * d / t_gemm / 1e6, 1)}
    return out


def run_extras_child():
    """Child of the default bench run (see run_ours): probes whose failure must not reach the parent."""
    import skin_sm3_b200 as sm3
    torch.cuda.set_device(0)
    out = {}
    for key, probe in (("small_shapes", small_shapes_probe), ("tc_kernels_cfg2", tc_kernel_probe),
                       ("kmeans", kmeans_probe), ("retrieval", retrieval_probe), ("projector_tail", projector_tail_probe),
                       ("head_block", head_block_probe)):
        try:
            out[key] = probe(sm3)
        except Exception as e:
            out[key] = {"error": repr(e)[:300]}
        print(json.dumps(out), flush=True)          # cumulative: the parent keeps the last complete line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS) + ["cfg3"], default="cfg4")
    ap.add_argument("--no-extras", action="store_true", help="skip cpu_baseline / cfg2 / heads (profiling runs)")
    ap.add_argument("--extras-child", action="store_true", help=argparse.SUPPRESS)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.extras_child:
        run_extras_child()
    elif args.impl == "reference":
        run_reference(args)
    elif args.workload == "cfg3":
        run_cfg3(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
