"""Sweep the tensor-core kernels' tuning knobs in ONE process on one GPU and print a JSON line per configuration:
K2 (forward) and K3 (backward) device-timed separately on resident inputs, at the shapes the reference trains at
(D = 128) and at cfg4.  Knobs: SM3_TC_FWD_BM (rows per forward CTA), SM3_TC_POLY (FMA-pipe exponentials per 8),
SM3_TC_GROUPS (softmax warp groups), SM3_TC_BWD_NS (S/H stages at D <= 128), SM3_TC_FWD_SPLITS / SM3_TC_BWD_SPLITS.

    python tools/tc_sweep.py [--out gpurun_out/tc_sweep.jsonl] [--quick]
"""
import argparse
import itertools
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import skin_sm3_b200 as sm3  # noqa: E402

KNOBS = ("SM3_TC_FWD_BM", "SM3_TC_POLY", "SM3_TC_GROUPS", "SM3_TC_BWD_NS", "SM3_TC_FWD_SPLITS", "SM3_TC_BWD_SPLITS")


def set_knobs(cfg):
    for k in KNOBS:
        os.environ.pop(k, None)
    for k, v in cfg.items():
        if v is not None:
            os.environ[k] = str(v)
    sm3.lib().sm3_debug_reload_env()


def ev_time(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); e1.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    shapes = [(4096, 128), (1024, 128), (256, 128)] + ([] if args.quick else [(32768, 256)])
    lines = []
    for n, d in shapes:
        T = 0.1
        g = torch.Generator(device="cuda").manual_seed(n)
        z, _ = sm3.core.normalize_pair(torch.randn(2 * n, d, generator=g, device="cuda"), None, torch.bfloat16)
        set_knobs({})
        pos, lse, nsum = sm3.core.stats_fwd(z, z, n, 0, n, T, sm3.ALGO_TC)
        _, gp, gl = sm3.core.loss(pos, lse, 1.0 / (2 * n))
        ref = (pos.clone(), nsum.clone())
        reps = 50 if n <= 4096 else 5
        fwd_cfgs = [dict(SM3_TC_FWD_BM=bm, SM3_TC_POLY=poly, SM3_TC_GROUPS=grp, SM3_TC_FWD_SPLITS=sp)
                    for bm, poly, grp, sp in itertools.product((256, 128), (0, 2), (1, 2), (None, 2, 4, 8, 16))
                    if not (bm == 256 and grp == 2)]          # the 256-row kernel has one softmax group
        bwd_cfgs = [dict(SM3_TC_BWD_NS=ns, SM3_TC_GROUPS=grp, SM3_TC_BWD_SPLITS=sp)
                    for ns, grp, sp in itertools.product((4, 2), (1, 2), (None, 1, 2, 3, 4, 8))]
        if args.quick:
            fwd_cfgs = [c for c in fwd_cfgs if c["SM3_TC_FWD_SPLITS"] is None]
            bwd_cfgs = [c for c in bwd_cfgs if c["SM3_TC_BWD_SPLITS"] is None]
        for kind, cfgs in (("fwd", fwd_cfgs), ("bwd", bwd_cfgs)):
            for cfg in cfgs:
                set_knobs(cfg)
                try:
                    if kind == "fwd":
                        p2, _, n2 = sm3.core.stats_fwd(z, z, n, 0, n, T, sm3.ALGO_TC)
                        ok = bool(torch.allclose(p2, ref[0], rtol=1e-5, atol=1e-6) and torch.allclose(n2, ref[1], rtol=1e-4))
                        ms = ev_time(lambda: sm3.core.stats_fwd(z, z, n, 0, n, T, sm3.ALGO_TC), reps)
                        flops = 2.0 * (2 * n) ** 2 * d
                    else:
                        ms = ev_time(lambda: sm3.core.stats_bwd(z, z, n, 0, n, T, gp, gl, nsum, gp, gl, nsum, sm3.ALGO_TC), reps)
                        ok, flops = True, 4.0 * (2 * n) ** 2 * d
                    rec = {"kernel": kind, "n_pairs": n, "dim": d, **{k: v for k, v in cfg.items() if v is not None},
                           "us": round(ms * 1e3, 2), "TFLOP/s": round(flops / ms / 1e9, 1), "ok": ok}
                except Exception as e:          # a knob combination a kernel does not support
                    rec = {"kernel": kind, "n_pairs": n, "dim": d, **{k: v for k, v in cfg.items() if v is not None},
                           "error": repr(e)[:160]}
                lines.append(rec)
                print(json.dumps(rec), flush=True)
    set_knobs({})
    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            for rec in lines:
                f.write(json.dumps(rec) + "\n")


if __name__ == "__main__":
    main()
