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
