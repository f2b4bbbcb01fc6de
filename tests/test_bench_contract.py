"""CPU: the measurement contract of bench.py that can be checked without a GPU -- the reference arm (`--impl
reference`: the reference's CPU algorithm on the host cores, bounded sample of the cfg-4 workload) prints ONE JSON
line with the contract's keys, and the flop model of the GPU arm reproduces the committed counter-derived figure."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    res = subprocess.run([sys.executable, "bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0"], cwd=ROOT,
                         capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "voigt_logL_evals_per_sec" and d["unit"] == "logL/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    from oracle import refshim
    assert d["cpu_baseline"]["kind"] == ("reference" if refshim.available() else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["gpu_launches"] == 0
    assert "cfg4" in d["config"]["workload"]


def test_flop_model_matches_the_committed_opcode_histogram():
    """bench.py's per-path flop model, fed with the path fractions of the committed bench line, must land within 10 % of
    the flops counted from the ncu opcode histogram of the same workload (profiles/ncu_opcodes_r02.json)."""
    sys.path.insert(0, ROOT)
    import bench
    line = json.load(open(os.path.join(ROOT, "profiles", "bench_r02_1gpu.json")))
    counters = json.load(open(os.path.join(ROOT, "profiles", "ncu_opcodes_r02.json")))
    r = line["roofline"]
    assert counters["samples_per_launch"] == line["config"]["batch_per_gpu"]
    assert abs(r["flop_per_logL_model"] / counters["fp32_flop_per_logL"] - 1.0) < 0.10
    # the constants are the ones the line was computed with: rebuild the model figure from the line's own fractions
    B, npix, E = line["config"]["batch_per_gpu"], 8192, r["evals_per_logL"]
    st = {"evals_total": E * B, "evals_far": r["frac_far"] * E * B, "evals_wing": r["frac_wing"] * E * B,
          "evals_mixed": r["frac_mixed"] * E * B, "evals_core": r["frac_core"] * E * B, "evals_core_precise": 0.0,
          "evals_core_straddle": 0.0}
    n = np.full(B, 10.0)                                   # mean LSF half-width of the cfg-4 prior: 8..14 pixels
    approx = bench.model_flops(st, {"nchunks": 32}, B, npix, n) / B
    assert abs(approx / r["flop_per_logL_model"] - 1.0) < 0.05


def test_committed_scaling_lines_are_one_consistent_set():
    """profiles/bench_r02_{1,2,4,8}gpu.json: every line carries the contract's keys, the strong records evaluate the
    same two global batches at every N, and their checksums (SHA-256 of the gathered logL bytes) are identical at
    N = 1, 2, 4, 8 -- the G-invariance evidence the documents quote."""
    lines = {n: json.load(open(os.path.join(ROOT, "profiles", "bench_r02_%dgpu.json" % n))) for n in (1, 2, 4, 8)}
    sums = {}
    for n, d in lines.items():
        assert d["n_gpus"] == n and d["metric"] == "voigt_logL_evals_per_sec" and d["scaling"] == "weak"
        for key in ("value", "ms_per_step", "e2e", "gpu_launches", "clocks", "roofline", "config", "dtype", "data"):
            assert key in d, (n, key)
        assert d["config"]["global_batch"] == n * d["config"]["batch_per_gpu"]
        assert d["e2e"]["value"] <= d["value"] * 1.001 and d["e2e"]["h2d_bytes_per_step"] > 0
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        for s in d["strong"]:
            assert s["bit_identical_to_single_gpu"] is True and s["batch_per_gpu"] * n >= s["global_batch"]
            sums.setdefault(s["global_batch"], set()).add(s["global_checksum"])
    assert set(sums) == {16384, 262144} and all(len(v) == 1 for v in sums.values()), sums
    # weak scaling of the committed set: no N loses more than 3 % per GPU against N = 1
    for n in (2, 4, 8):
        assert lines[n]["value"] >= 0.97 * n * lines[1]["value"]
