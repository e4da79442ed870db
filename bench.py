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
# our kernels per step on the production path (one GPU: l2norm_fwd, infonce_fwd, finalize+loss, bwd_prep, infonce_bwd,
# l2norm_bwd; multi-rank fused exchange: l2norm+scatter, infonce_fwd, loss+stats scatter, infonce_bwd, l2norm_bwd)
KERNELS_PER_STEP = {True: 6, False: 5}


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


def synth(n_global, d, rank, world, device):
    """p1, p2 ~ N(0,1) (the projector ends in an affine-free BatchNorm, simclr.py:26), bf16, this rank's pairs."""
    n_local = n_global // world
    g = torch.Generator(device="cpu").manual_seed(SEED + rank)
    p1 = torch.randn(n_local, d, generator=g).bfloat16()
    p2 = torch.randn(n_local, d, generator=g).bfloat16()
    return p1, p2


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

    if world == 1:
        host = sm3.HostInfoNCE(n_local, d, torch.bfloat16, sm3.ALGO_AUTO)             # sm3_infonce_host: synchronises
        pipe = sm3.HostInfoNCEPipeline(n_local, d, torch.bfloat16, sm3.ALGO_AUTO, depth)  # sm3_host_pipe_*
        sink = []

        def run_sync(k=steps):
            for _ in range(k):
                host(hp1, hp2, T)

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


def cpu_reference(n, d, T, budget_s=20.0, max_reps=8):
    """The reference's materialising CPU path (oracle/ref_port.py, an op-for-op torch port pinned to the real
    reference by tests/test_oracle_golden.py) on the host cores.  Bounded sample -> pairs/s."""
    from oracle import ref_port
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(SEED)
    p1 = torch.randn(n, d, generator=g)
    p2 = torch.randn(n, d, generator=g)
    m = 2 * n
    full = m * m * 4 * 8 < 24e9                    # ~8 live [M,M] fp32 temporaries must fit comfortably
    if full:
        fn = lambda: ref_port.port_infonce_step(p1, p2, T)           # noqa: E731
        scale, sample = 1.0, f"full step N={n} D={d} fp32 (reference op sequence, {cores} threads)"
    else:
        rows = 512
        fn = lambda: ref_port.port_infonce_step_rowblock(p1, p2, T, 1024, rows)   # noqa: E731
        scale = m / rows
        sample = (f"row block of {rows}/{m} logits rows x all {m} columns of N={n} D={d} fp32 "
                  f"(full [M,M] step needs ~{m * m * 4 * 6 / 1e9:.0f} GB); time scaled x{scale:.0f}")
    fn()
    times = []
    t_start = time.perf_counter()
    while len(times) < max_reps and (time.perf_counter() - t_start) < budget_s:
        t0 = time.perf_counter()
        fn()
        times.append(time.perf_counter() - t0)
    t = statistics.median(times) * scale
    return dict(value=n / t, unit="pairs/s", cores=cores, kind="port", sample=sample, ms_per_step=t * 1e3,
                reps=len(times))


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
    p1, p2 = synth(n, d, rank, world, "cuda")
    flush = torch.zeros(64 * 1024 * 1024, device="cuda")            # 256 MB > 126 MB L2
    with ClockSampler(local) as clk:
        total_ms, stages, loss = time_device(sm3, p1, p2, T, group, world, args.steps, args.warmup, flush)
        # `value` is measured on the production path (the whole step enqueued by ONE C call, no per-stage event marks, no
        # Python between the kernels); the profiled pass above only supplies the per-kernel durations for the roofline.
        total_ms, _, loss = time_device(sm3, p1, p2, T, group, world, args.steps, args.warmup, flush, profile=False)
    e2e_ms, e2e_sync_ms, h2d, d2h = time_e2e(sm3, p1, p2, T, group, world, args.steps, args.warmup)
    ms_step = total_ms / args.steps
    m_cols, m_rows = 2 * n, 2 * n // world
    comm_used = "none"
    if world > 1:
        from skin_sm3_b200 import peer as _peer
        comm_used = "NCCL all-gather"
        if _peer._CACHE:
            mc = any(b.multicast for b in _peer._CACHE.values())
            comm_used = "NVLink peer memory (symmetric memory" + (", NVSwitch multicast stores)" if mc else ", unicast stores)")
            if os.environ.get("SM3_PEER_FUSED", "1") != "0" and (n // world) % 128 == 0:
                comm_used += ", fused exchange: scatter+signal in the producer kernels, waits inside K2/K3"
    flops_bwd = 4.0 * m_rows * m_cols * d
    flops_fwd = 2.0 * m_rows * m_cols * d
    line = {
        "metric": "contrastive pairs/sec (fwd+bwd)", "value": n / (ms_step * 1e-3), "unit": "pairs/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": wl["name"], "global_pairs": n, "rows_per_rank": m_rows, "dim": d, "temperature": T,
                   "terms": 1, "sharding": (f"row-block x{world}, global negatives exchanged over " + comm_used) if world > 1 else "single GPU",
                   "l2": "256 MB flush write between timed steps (inputs are L2-sized by design)"},
        "loss": loss,
        "e2e": {"value": n / (e2e_ms / args.steps * 1e-3), "unit": "pairs/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps,
                "sync_value": n / (e2e_sync_ms / args.steps * 1e-3), "sync_ms_per_step": e2e_sync_ms / args.steps,
                "api": ("sm3_host_pipe_submit/wait (C ABI, pinned host buffers, 2-slot copy/compute pipeline); "
                        "sync_value = one synchronous sm3_infonce_host call per step") if world == 1 else
                       ("skin_sm3_b200.fused_infonce(group=WORLD) fed from / drained to pinned host memory on two copy "
                        "streams, 2 slots; sync_value = drained after every step")},
        "gpu_launches": KERNELS_PER_STEP[world == 1] * args.steps,
        "stages_ms": {k: round(v, 4) for k, v in stages.items()},
        "clocks": clk.summary(),
    }
    t_bwd = stages.get("stats_bwd")
    t_fwd = stages.get("stats_fwd")
    if t_bwd:
        ach = flops_bwd / (t_bwd * 1e-3) / 1e12
        line["roofline"] = {"bound": "tensor", "kernel": "infonce_tc_bwd_kernel (K3; includes its 1-block prep kernel)",
                            "achieved": ach, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": ach / pk["tflops"],
                            "traffic": (ncu_traffic().get("infonce_tc_bwd_kernel", {}).get("dram_bytes_per_launch")
                                        if (args.workload == "cfg4" and world == 1) else None),
                            "traffic_source": ncu_traffic().get("source"),
                            "peak_source": pk["source"] + ", burst bf16", "flops_per_launch": flops_bwd}
    if t_fwd:
        ach = flops_fwd / (t_fwd * 1e-3) / 1e12
        line["roofline_fwd"] = {"bound": "tensor", "kernel": "infonce_tc_fwd_kernel (K2; includes the finalize kernel)",
                                "achieved": ach, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": ach / pk["tflops"]}
    line["step_tc_frac"] = (flops_fwd + flops_bwd) / (ms_step * 1e-3) / 1e12 / pk["tflops"]
    if world == 1:
        try:
            gms = time_graph(sm3, p1, p2, T, args.steps, args.warmup, flush)
            line["cuda_graph"] = {"ms_per_step": gms, "value": n / (gms * 1e-3), "unit": "pairs/s",
                                  "what": "same step replayed from a CUDA graph (GraphedInfoNCE), device-timed"}
        except Exception as e:
            line["cuda_graph"] = {"error": repr(e)[:300]}
    if rank == 0 and world == 1 and not args.no_extras:
        line["cpu_baseline"] = cpu_reference(n, d, T)
        if args.workload != "cfg2":     # configs[1] rides along in the same line
            w2 = WORKLOADS["cfg2"]
            q1, q2 = synth(w2["n"], w2["d"], 0, 1, "cuda")
            _, st2, _ = time_device(sm3, q1, q2, w2["T"], None, 1, args.steps, args.warmup, flush)          # per-stage events
            ms2, _, _ = time_device(sm3, q1, q2, w2["T"], None, 1, args.steps, args.warmup, flush, profile=False)
            e2, e2s, _, _ = time_e2e(sm3, q1, q2, w2["T"], None, 1, args.steps, args.warmup)
            try:
                g2 = time_graph(sm3, q1, q2, w2["T"], args.steps, args.warmup, flush)
            except Exception as e:
                g2 = None
                print(f"[bench] cfg2 graph replay failed: {e!r}", file=sys.stderr)
            f2 = 6.0 * (2 * w2["n"]) ** 2 * w2["d"]
            line["cfg2"] = {"workload": w2["name"], "value": w2["n"] / (ms2 / args.steps * 1e-3), "unit": "pairs/s",
                            "ms_per_step": ms2 / args.steps, "e2e_value": w2["n"] / (e2 / args.steps * 1e-3),
                            "e2e_sync_value": w2["n"] / (e2s / args.steps * 1e-3),
                            "cuda_graph_value": (w2["n"] / (g2 * 1e-3)) if g2 else None,
                            "cuda_graph_ms_per_step": g2,
                            "step_tc_frac": f2 / (ms2 / args.steps * 1e-3) / 1e12 / pk["tflops"],
                            "stages_ms": {k: round(v, 4) for k, v in st2.items()},
                            "cpu_baseline": cpu_reference(w2["n"], w2["d"], w2["T"], budget_s=12.0, max_reps=4)}
            try:
                line["cfg2"]["gpu_reference_port"] = gpu_reference_port(w2["n"], w2["d"], w2["T"])
            except Exception as e:
                line["cfg2"]["gpu_reference_port"] = {"error": repr(e)[:200]}
        try:
            line["cfg4_sweep"] = cfg4_sweep(sm3, args.steps, flush)
        except Exception as e:
            line["cfg4_sweep"] = {"error": repr(e)[:300]}
        try:
            line["heads"] = heads_probe(sm3, pk)
        except Exception as e:   # the head probe must never take the headline down
            line["heads"] = {"error": repr(e)}
        # The newest probes run in a CHILD process with a timeout: whatever happens there (exception, device trap, hang)
        # cannot touch the headline numbers above or keep this process from printing its line.
        import subprocess
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--extras-child"], capture_output=True,
                               text=True, timeout=120)
            got = _last_json(r.stdout)
            line.update(got if got is not None else {"extras_child": {"error": (r.stderr or "no output")[-300:]}})
        except subprocess.TimeoutExpired as e:
            got = _last_json(e.stdout.decode() if isinstance(e.stdout, bytes) else e.stdout)
            line.update(got or {})
            line["extras_child"] = {"error": "timeout after 120 s"}
        except Exception as e:
            line["extras_child"] = {"error": repr(e)[:300]}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


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
    """K2 and K3 timed on their own at cfg2 (4096 x 128, the D = 128 regime of every reference config), for the knob values
    that have been parity-checked on B200: forward rows per CTA 256 / 128 and FMA-pipe exponentials 0 / 2 per 8; backward
    S/H stages 4 / 2.  Feeds DESIGN.md section 9 item 1.  Microseconds per launch (finalize / prep kernels included)."""
    n, d, T = 4096, 128, 0.1
    z, _ = sm3.core.normalize_pair(torch.randn(2 * n, d, device="cuda"), None, torch.bfloat16)
    pos, lse, nsum = sm3.core.stats_fwd(z, z, n, 0, n, T, sm3.ALGO_TC)
    _, gp, gl = sm3.core.loss(pos, lse, 1.0 / (2 * n))
    knobs = ("SM3_TC_FWD_BM", "SM3_TC_POLY", "SM3_TC_BWD_NS")
    saved = {k: os.environ.get(k) for k in knobs}
    out = {}
    try:
        for bm, poly in (("256", "0"), ("256", "2"), ("128", "0")):      # the combinations that have run on B200 before
            os.environ["SM3_TC_FWD_BM"], os.environ["SM3_TC_POLY"] = bm, poly
            sm3.lib().sm3_debug_reload_env()
            out[f"fwd_bm{bm}_poly{poly}_us"] = round(_ev_us(lambda: sm3.core.stats_fwd(z, z, n, 0, n, T, sm3.ALGO_TC)), 1)
        os.environ.pop("SM3_TC_FWD_BM", None); os.environ.pop("SM3_TC_POLY", None)
        for ns in ("4", "2"):
            os.environ["SM3_TC_BWD_NS"] = ns
            sm3.lib().sm3_debug_reload_env()
            out[f"bwd_stages{ns}_us"] = round(_ev_us(
                lambda: sm3.core.stats_bwd(z, z, n, 0, n, T, gp, gl, nsum, gp, gl, nsum, sm3.ALGO_TC)), 1)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        sm3.lib().sm3_debug_reload_env()
    return out


def tc_experimental_probe(sm3, emit):
    """Knob combinations that have NOT been run on a B200 yet (two softmax groups with 4 backward stages, FMA-pipe
    exponentials in the 128-row forward, forced split counts), at the D = 128 shapes.  Only ever called from the
    extras child: each result is checked against the default configuration's output ("ok") and emitted immediately,
    so a trap in one combination costs the combinations after it, nothing else."""
    knobs = ("SM3_TC_FWD_BM", "SM3_TC_POLY", "SM3_TC_GROUPS", "SM3_TC_BWD_NS", "SM3_TC_FWD_SPLITS", "SM3_TC_BWD_SPLITS")

    def set_knobs(cfg):
        for k in knobs:
            os.environ.pop(k, None)
        os.environ.update({k: str(v) for k, v in cfg.items()})
        sm3.lib().sm3_debug_reload_env()

    out = []
    for n in (4096, 1024, 256):
        d, T = 128, 0.1
        z, _ = sm3.core.normalize_pair(torch.randn(2 * n, d, device="cuda"), None, torch.bfloat16)
        set_knobs({})
        pos, lse, nsum = sm3.core.stats_fwd(z, z, n, 0, n, T, sm3.ALGO_TC)
        _, gp, gl = sm3.core.loss(pos, lse, 1.0 / (2 * n))
        ws, npart = sm3.core.stats_bwd(z, z, n, 0, n, T, gp, gl, nsum, gp, gl, nsum, sm3.ALGO_TC)
        dz_ref = sm3.core.sum_partials(ws, npart, 2 * n, d).clone()
        fwd = [{"SM3_TC_FWD_BM": 128, "SM3_TC_POLY": 2}, {"SM3_TC_FWD_BM": 128, "SM3_TC_GROUPS": 2},
               {"SM3_TC_FWD_BM": 128, "SM3_TC_GROUPS": 2, "SM3_TC_POLY": 2},
               {"SM3_TC_FWD_SPLITS": 2}, {"SM3_TC_FWD_SPLITS": 8}, {"SM3_TC_FWD_BM": 128, "SM3_TC_FWD_SPLITS": 8}]
        bwd = [{"SM3_TC_GROUPS": 2, "SM3_TC_BWD_NS": 4}, {"SM3_TC_GROUPS": 2, "SM3_TC_BWD_NS": 2},
               {"SM3_TC_BWD_SPLITS": 1}, {"SM3_TC_BWD_SPLITS": 3}, {"SM3_TC_BWD_SPLITS": 4},
               {"SM3_TC_BWD_SPLITS": 4, "SM3_TC_GROUPS": 2}]
        for kind, cfgs in (("fwd", fwd), ("bwd", bwd)):
            for cfg in cfgs:
                rec = {"kernel": kind, "pairs": n, **cfg}
                try:
                    set_knobs(cfg)
                    if kind == "fwd":
                        p2, _, n2 = sm3.core.stats_fwd(z, z, n, 0, n, T, sm3.ALGO_TC)
                        rec["ok"] = bool(torch.allclose(p2, pos, rtol=1e-5, atol=1e-6) and torch.allclose(n2, nsum, rtol=1e-3))
                        rec["us"] = round(_ev_us(lambda: sm3.core.stats_fwd(z, z, n, 0, n, T, sm3.ALGO_TC)), 1)
                    else:
                        w2, np2 = sm3.core.stats_bwd(z, z, n, 0, n, T, gp, gl, nsum, gp, gl, nsum, sm3.ALGO_TC)
                        dz = sm3.core.sum_partials(w2, np2, 2 * n, d)
                        rec["ok"] = bool((dz - dz_ref).abs().max().item() <= 2e-2 * dz_ref.abs().max().item())
                        rec["us"] = round(_ev_us(
                            lambda: sm3.core.stats_bwd(z, z, n, 0, n, T, gp, gl, nsum, gp, gl, nsum, sm3.ALGO_TC)), 1)
                except Exception as e:
                    rec["error"] = repr(e)[:200]
                out.append(rec)
                emit(out)
    set_knobs({})
    return out


def hbm_experimental_probe(sm3, emit):
    """Opt-in variants of the HBM-bound kernels that have not run on hardware yet (extras child only): the persistent
    normalise-backward (SM3_K1_BWD_VARIANT=1) and the deep-prefetch BCE kernels (SM3_BCE_VARIANT=2|3), each checked
    against the default's output and timed at the bandwidth-sized shapes of `heads`."""
    out = []
    pk = peaks()

    def rec(name, variant, ms, nbytes, ok):
        out.append({"kernel": name, "variant": variant, "us": round(ms * 1e3, 2), "frac_hbm": round(nbytes / ms / 1e6 / pk["hbm"], 3),
                    "ok": bool(ok)})
        emit(out)

    try:
        for M, D, parts in ((1 << 21, 256, 1), (65536, 256, 2)):
            z, inv = sm3.core.normalize_pair(torch.randn(M, D, device="cuda", dtype=torch.bfloat16), None, torch.bfloat16)
            dz = torch.randn(parts, M, D, device="cuda", dtype=torch.float32)
            nbytes = M * D * (4 * parts + 2 + 2) + 4 * M
            ref = None
            for v in ("0", "1"):
                os.environ["SM3_K1_BWD_VARIANT"] = v
                o, _ = sm3.core.normalize_bwd(dz, parts, 1.0, z, inv, M, 0, torch.bfloat16)
                ref = o.clone() if ref is None else ref
                us = _ev_us(lambda: sm3.core.normalize_bwd(dz, parts, 1.0, z, inv, M, 0, torch.bfloat16), reps=10)
                rec(f"l2norm_bwd_{M}x{D}_p{parts}", v, us / 1e3, nbytes, torch.equal(o, ref))
            del z, inv, dz, ref
    except Exception as e:
        out.append({"kernel": "l2norm_bwd", "error": repr(e)[:200]}); emit(out)
    finally:
        os.environ.pop("SM3_K1_BWD_VARIANT", None)
    try:
        B = 1 << 22
        x = torch.randn(B, 24, device="cuda", dtype=torch.bfloat16, requires_grad=True)
        t = (torch.rand(B, 24, device="cuda") < 0.3).to(torch.bfloat16)
        ref = None
        for v in ("1", "2", "3"):
            os.environ["SM3_BCE_VARIANT"] = v
            x.grad = None
            loss = sm3.bce_with_logits(x, t)
            loss.backward()
            cur = (float(loss.detach()), x.grad.float().clone())
            ref = cur if ref is None else ref
            ok = abs(cur[0] - ref[0]) <= 1e-5 * abs(ref[0]) and \
                (cur[1] - ref[1]).abs().max().item() <= 1e-2 * ref[1].abs().max().item()
            us = _ev_us(lambda: sm3.bce_with_logits(x, t), reps=20)
            rec("bce_b4M", v, us / 1e3, B * 24 * 6, ok)
    except Exception as e:
        out.append({"kernel": "bce", "error": repr(e)[:200]}); emit(out)
    finally:
        os.environ.pop("SM3_BCE_VARIANT", None)
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
    """Reference arm: the reference's own CPU implementation of the path (op-for-op port in oracle/ref_port.py;
    the Python reference itself cannot travel to the GPU box) on all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    wl = WORKLOADS["cfg4" if args.workload == "cfg3" else args.workload]
    n, d, T = wl["n"], wl["d"], wl["T"]
    reps = max(1, args.steps)
    r = cpu_reference(n, d, T, budget_s=60.0, max_reps=min(reps, 8))
    line = {
        "impl": "reference", "metric": "contrastive pairs/sec (fwd+bwd)", "value": r["value"], "unit": "pairs/s",
        "n_gpus": world, "steps": r["reps"], "warmup": 1, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "global_pairs": n, "dim": d, "temperature": T, "terms": 1,
                   "device": "host CPU"},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_extras_child():
    """Child of the default bench run (see run_ours): probes whose failure must not reach the parent."""
    import skin_sm3_b200 as sm3
    torch.cuda.set_device(0)
    out = {}
    for key, probe in (("small_shapes", small_shapes_probe), ("tc_kernels_cfg2", tc_kernel_probe),
                       ("kmeans", kmeans_probe)):
        try:
            out[key] = probe(sm3)
        except Exception as e:
            out[key] = {"error": repr(e)[:300]}
        print(json.dumps(out), flush=True)          # cumulative: the parent keeps the last complete line

    # last: code that has never run on the hardware, each result checked against the default's and emitted at once
    for key, probe in (("hbm_experimental", hbm_experimental_probe), ("tc_experimental", tc_experimental_probe)):
        def emit(partial, key=key):
            out[key] = partial
            print(json.dumps(out), flush=True)
        try:
            probe(sm3, emit)
        except Exception as e:
            out[key + "_error"] = repr(e)[:300]
            print(json.dumps(out), flush=True)


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
