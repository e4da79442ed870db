"""Time every variant of the HBM-bound kernels (K1 fwd, K4, K5) at a bandwidth-sized shape on one GPU, in ONE process
(the variant selectors SM3_K1_FWD_VARIANT / SM3_CE_VARIANT / SM3_BCE_VARIANT are read at every launch), and print one
JSON line per (kernel, variant): us per launch, algorithmic GB/s (DESIGN.md section 4 byte counts), fraction of the
measured HBM copy rate.  Outputs of the variants are compared with variant 0 (bitwise for K1, 1e-6 relative for the
losses) so a faster-but-wrong variant cannot be picked.

    python tools/hbm_variants.py [--out gpurun_out/hbm_variants.jsonl]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import skin_sm3_b200 as sm3  # noqa: E402


def hbm_peak():
    try:
        with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"])
    except Exception:
        return 6650.0


def ev_time(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):                                   # best of 3 batches of `reps` back-to-back launches
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
    args = ap.parse_args()
    peak = hbm_peak()
    lines = []

    def emit(kernel, variant, ms, nbytes, ok, extra=None):
        d = {"kernel": kernel, "variant": variant, "us": round(ms * 1e3, 2), "GB/s": round(nbytes / ms / 1e6, 1),
             "frac_hbm": round(nbytes / ms / 1e6 / peak, 3), "matches_variant0": bool(ok)}
        if extra:
            d.update(extra)
        lines.append(d)
        print(json.dumps(d), flush=True)

    # ---- K1 forward: 2M x 256 (bf16 -> bf16) and (fp32 -> bf16, what fp32 projector outputs give) ----
    M, D = 1 << 21, 256
    for in_dt, tag in ((torch.bfloat16, "bf16->bf16"), (torch.float32, "fp32->bf16")):
        p = torch.randn(M, D, device="cuda", dtype=in_dt)
        nbytes = M * D * (p.element_size() + 2) + 4 * M
        ref = None
        for v in ("0", "1", "2"):
            os.environ["SM3_K1_FWD_VARIANT"] = v
            z, inv = sm3.core.normalize_pair(p, None, torch.bfloat16)
            if ref is None:
                ref = (z.clone(), inv.clone())
            ok = torch.equal(z, ref[0]) and torch.equal(inv, ref[1])
            ms = ev_time(lambda: sm3.core.normalize_pair(p, None, torch.bfloat16))
            emit(f"l2norm_fwd_2Mx256 {tag}", v, ms, nbytes, ok)
        del p, z, inv, ref
    os.environ.pop("SM3_K1_FWD_VARIANT", None)
    # the real cfg4 shape (L2 resident): 65536 x 256
    p = torch.randn(65536, 256, device="cuda", dtype=torch.bfloat16)
    for v in ("0", "1", "2"):
        os.environ["SM3_K1_FWD_VARIANT"] = v
        ms = ev_time(lambda: sm3.core.normalize_pair(p, None, torch.bfloat16), reps=50)
        emit("l2norm_fwd_65536x256 bf16->bf16 (L2 resident)", v, ms, 65536 * 256 * 4 + 4 * 65536, True)
    os.environ.pop("SM3_K1_FWD_VARIANT", None)
    del p

    # ---- K1 backward: default vs the experimental persistent form ----
    M, D = 1 << 21, 256
    z, inv = sm3.core.normalize_pair(torch.randn(M, D, device="cuda", dtype=torch.bfloat16), None, torch.bfloat16)
    dz = torch.randn(M, D, device="cuda", dtype=torch.float32)
    ref = None
    for v in ("0", "1"):
        os.environ["SM3_K1_BWD_VARIANT"] = v
        o, _ = sm3.core.normalize_bwd(dz, 1, 1.0, z, inv, M, 0, torch.bfloat16)
        ref = o.clone() if ref is None else ref
        ms = ev_time(lambda: sm3.core.normalize_bwd(dz, 1, 1.0, z, inv, M, 0, torch.bfloat16), reps=10)
        emit("l2norm_bwd_2Mx256 bf16", v, ms, M * D * 8 + 4 * M, torch.equal(o, ref))
    os.environ.pop("SM3_K1_BWD_VARIANT", None)
    del z, inv, dz, ref

    # ---- K4 / K5 at B = 4M rows x 24 logits, and the reference-sized B = 512 ----
    for B in (1 << 22, 512):
        x = torch.randn(B, 24, device="cuda", dtype=torch.bfloat16, requires_grad=True)
        y = torch.stack([torch.randint(0, c, (B,), device="cuda") for c in sm3.NUM_CLASSES], 1)
        t = torch.nn.functional.one_hot(y[:, 0], 24).to(torch.bfloat16)
        ref = None
        for v in ("0", "1"):
            os.environ["SM3_CE_VARIANT"] = v
            x.grad = None
            loss = sm3.multihead_ce(x, y)
            loss.backward()
            cur = (float(loss), x.grad.clone())
            if ref is None:
                ref = cur
            ok = abs(cur[0] - ref[0]) <= 1e-6 * abs(ref[0]) and torch.equal(cur[1], ref[1])
            ms = ev_time(lambda: sm3.multihead_ce(x, y))
            emit(f"ce8_b{B}", v, ms, B * 24 * 4 + B * 64 + 4, ok)
        os.environ.pop("SM3_CE_VARIANT", None)
        ref = None
        for v in ("0", "1", "2", "3"):                  # 2 / 3: experimental deep-prefetch kernels
            os.environ["SM3_BCE_VARIANT"] = v
            x.grad = None
            loss = sm3.bce_with_logits(x, t)
            loss.backward()
            cur = (float(loss), x.grad.clone())
            if ref is None:
                ref = cur
            ok = abs(cur[0] - ref[0]) <= 2e-6 * abs(ref[0]) and \
                (cur[1].float() - ref[1].float()).abs().max().item() <= 1e-2 * ref[1].float().abs().max().item()
            ms = ev_time(lambda: sm3.bce_with_logits(x, t))
            emit(f"bce_b{B}", v, ms, B * 24 * 6, ok, {"loss": cur[0]})
        os.environ.pop("SM3_BCE_VARIANT", None)
        del x, y, t

    if args.out:
        os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
        with open(args.out, "w") as f:
            for d in lines:
                f.write(json.dumps(d) + "\n")


if __name__ == "__main__":
    main()
