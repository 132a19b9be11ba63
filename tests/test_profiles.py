"""Measurement hygiene (no GPU needed): the DRAM-traffic constants bench.py reports (profiles/traffic.json) are exactly what the
committed ncu summaries say, and carry the commit they were captured at."""
import json
import os
import sys

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "scripts"))


def test_traffic_json_equals_the_ncu_summaries_it_cites():
    import make_traffic
    d = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    assert set(make_traffic.SOURCES) <= set(d)
    for key, files in make_traffic.SOURCES.items():
        e = d[key]
        assert e["captured_at"], key
        assert e["source"] == [f"profiles/{f}" for f in files]
        want = sum(make_traffic.dram_bytes(os.path.join(ROOT, "profiles", f)) for f in files)
        assert e["bytes"] == want, (key, e["bytes"], want)
    # the algorithmic floor of one 16384^2 FP32 launch is 16 B x 268,435,456 cells: the captures sit within 15 % above it
    floor = 16 * 16384 * 16384
    assert floor <= d["fast_f32_16384"]["bytes"] <= 1.15 * floor
    assert floor <= d["fast2_f32_16384"]["bytes"] <= 1.15 * floor


def test_bench_reads_traffic_entries():
    import bench
    b, at, src = bench.traffic_entry("fast_f32_16384")
    assert b and at and src


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the reference TU on the host cores; no GPU involved) prints ONE JSON line with the contract's keys,
    the same metric / unit / config as the native arm, and `e2e` equal to its own value with zero copy bytes."""
    import subprocess
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-n", "256"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "Gcell-updates/s" and d["unit"] == "Gcell/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Gcell/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "16384x16384" in d["config"]["workload"]
