"""bench.py contract checks that run without a GPU: the reference arm prints exactly one JSON line with the keys the
driver reads, and the "ours" arm refuses to run without CUDA instead of falling back."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg2",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "pairs/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["steps"] == 1 and d["warmup"] == 0             # the arm runs exactly the steps it was asked for
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_runs_the_reference_sized_case_in_full():
    """cfg1 (batch 64 x dim 128, BASELINE.json configs[0]) is small enough that the reference arm times the FULL
    materialising step, not a row-block sample."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1",
                        "--steps", "3"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([l for l in r.stdout.splitlines() if l.strip()][-1])
    assert d["config"]["global_pairs"] == 64 and d["config"]["dim"] == 128
    assert d["cpu_baseline"]["sample"].startswith("full step") and d["estimated"] is False
    # the real reference (vendored by oracle/vendor_ref.py) is what runs when it is there
    import os as _os
    if _os.path.isdir(_os.path.join(ROOT, "oracle", "_ref", "skin_sm3")):
        assert d["cpu_baseline"]["kind"] == "reference"
    # same `config` object as our arm prints for this workload (the driver compares them)
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", _os.path.join(ROOT, "bench.py"))
    bm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bm)
    assert d["config"] == bm.config_of(bm.WORKLOADS["cfg1"], 64, 128, 0.1, 1)


def test_ours_arm_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--no-extras"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and not r.stdout.strip().startswith("{")


def test_committed_bench_line_carries_the_contract_keys():
    """The recorded B200 line (profiles/r01_bench_final.json) has every key the bench contract names, the roofline is
    self-consistent (frac = achieved / peak, achieved = algorithmic flops / the measured stage time) and the end-to-end
    figure really moved its bytes."""
    d = json.load(open(os.path.join(ROOT, "profiles", "r01_bench_final.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline"):
        assert k in d, k
    assert d["unit"] == "pairs/s" and d["n_gpus"] == 1 and d["warmup"] >= 3 and d["gpu_launches"] > 0
    n = d["config"]["global_pairs"]
    assert abs(d["value"] - n / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    r = d["roofline"]
    assert r["bound"] == "tensor" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    m = 2 * n
    assert r["flops_per_launch"] == 4.0 * m * m * d["config"]["dim"]
    assert abs(r["achieved"] - r["flops_per_launch"] / (d["stages_ms"]["stats_bwd"] * 1e-3) / 1e12) < 0.01 * r["achieved"]
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 2 * n * d["config"]["dim"] * 2 and e["d2h_bytes_per_step"] == e["h2d_bytes_per_step"] + 4
    assert 0 < e["value"] <= d["value"] * 1.05 and e["value"] != d["value"]
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and "sample" in c
    bad = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert not bad & set(d["clocks"]["reasons"])


def test_round2_bench_line_carries_parity_and_consistent_roofline():
    """The recorded round-2 B200 line (profiles/r02_bench_final.json): contract keys, a parity block measured on the
    production path, the real reference as CPU baseline (marked estimated where it is), the dominant-kernel roofline
    consistent with the stage time taken inside the production call, the forward's executed-vs-algorithmic FLOPs when the
    symmetric kernel ran, and the next-row timings at a level the driver's record keeps."""
    d = json.load(open(os.path.join(ROOT, "profiles", "r02_bench_final.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline", "cpu_baseline", "parity"):
        assert k in d, k
    n, dim = d["config"]["global_pairs"], d["config"]["dim"]
    assert d["unit"] == "pairs/s" and d["n_gpus"] == 1 and d["warmup"] >= 3 and d["gpu_launches"] > 0 and d["vs_baseline"] is None
    assert abs(d["value"] - n / (d["ms_per_step"] * 1e-3)) < 1e-6 * d["value"]
    r = d["roofline"]
    m = 2 * n
    assert r["bound"] == "tensor" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert r["flops_per_launch"] == 4.0 * m * m * dim
    assert abs(r["achieved"] - r["flops_per_launch"] / (r["stages_ms"]["infonce_bwd"] * 1e-3) / 1e12) < 0.01 * r["achieved"]
    assert abs(r["launch_ms"] - r["stages_ms"]["infonce_bwd"]) < 1e-3 and "production path" in r["stage_source"]
    f = r["fwd"]
    assert f["flops_per_launch"] == 2.0 * m * m * dim and 0 < f["executed_flops_per_launch"] <= f["flops_per_launch"]
    if "fwdsym" in f["kernel"]:
        assert f["executed_flops_per_launch"] < 0.55 * f["flops_per_launch"]
    p = r["parity"]
    assert p["ok"] is True and p["loss_relerr"] < 1e-4 and p["grad_relerr_rowblock"] < p["tolerance"] <= 2e-2 and p["rows_checked"] >= 256
    c = d["cpu_baseline"]
    assert c["kind"] == "reference" and c["cores"] >= 1 and c["value"] > 0 and c["estimated"] is True and c["scale_factor"] > 1
    e = d["e2e"]
    assert e["h2d_bytes_per_step"] == 2 * n * dim * 2 and e["d2h_bytes_per_step"] == e["h2d_bytes_per_step"] + 4
    assert 0 < e["value"] <= d["value"] * 1.05 and e["value"] != d["value"]
    for key in ("cfg2", "cfg4_sweep", "hbm_kernels", "retrieval", "kmeans", "projector_tail", "small_shapes"):
        assert key in r, key
    c2 = r["cfg2"]
    assert c2["parity"]["grad_relerr_all_rows"] < 2e-2 and c2["cpu_baseline"]["kind"] == "reference"
    bad = {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    assert not bad & set(d["clocks"]["reasons"])
