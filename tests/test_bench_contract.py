"""CPU: the parts of bench.py's contract that need no GPU -- the reference arm's JSON line (keys, units, the zero-copy
`e2e` object, the `cpu_baseline` description) and the roofline arithmetic helpers."""
import json
import os
import subprocess
import sys

from helpers import ROOT


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="4")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--batch", "64"],
                       capture_output=True, text=True, cwd=ROOT, env=env, timeout=600)
    assert p.returncode == 0, p.stderr[-500:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines                                    # exactly one JSON line on stdout
    d = json.loads(lines[0])
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert d["impl"] == "reference" and d["metric"] == base["metric"] and d["unit"] == "patients/sec"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] >= 1 and d["ms_per_step"] > 0 and d["value"] > 0
    assert "configs[1]" in d["config"]["workload"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_roofline_arithmetic():
    sys.path.insert(0, ROOT)
    import bench
    # 1.455 GFLOP forward per patient at L48/N16/I49, fwd+bwd = 3x
    assert abs(bench.total_flops_per_patient() / 3 - 1.455e9) / 1.455e9 < 0.01
    fl_tn, fl_wg = bench.gemm_flops_per_step(512)
    assert fl_wg * 2 == fl_tn and 0.95 < (fl_tn + fl_wg) / (512 * bench.total_flops_per_patient()) < 1.0   # GEMMs = 97.8 % of the FLOPs
    # profiles/r1_step_bytes.md: 7.87 GB of algorithmic GEMM traffic per 512-patient step, linear in the batch
    assert abs(bench.gemm_bytes_per_step(512) - 7.88e9) / 7.88e9 < 0.01
    assert bench.gemm_bytes_per_step(1024) == 2 * bench.gemm_bytes_per_step(512)
    burst, sust, hbm, src = bench.peaks()
    assert 1000 < sust <= burst < 2500 and 5000 < hbm < 8000 and src.split()[0] in ("measured", "fallback")
    # the other BASELINE configs are selectable and change the algorithmic FLOPs
    bench.WL.update(bench.WORKLOADS["inspect"])
    try:
        assert abs(bench.total_flops_per_patient() / 3 - 12.085e9) / 12.085e9 < 0.01
    finally:
        bench.WL.update(bench.WORKLOADS["pheno512"])
