"""profiles/traffic.json FROM the committed ncu summaries (so that the two cannot disagree; tests/test_profiles.py re-derives it).

    python scripts/make_traffic.py [commit]

bench.py reports `roofline.traffic` = dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, taken from
the `ncu --set full` capture summarised under profiles/ (a pair = far pass + general pass)."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROF = os.path.join(ROOT, "profiles")
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
# key -> summaries whose DRAM bytes add up to one launch (pair)
SOURCES = {
    "fast_f32_16384": ["r02_ncu_seeded_summary.md"],
    "fast_f32_16384_dense": ["r02_ncu_dense_summary.md"],
    "fast2_f32_16384": ["r02_ncu_far2_summary.md", "r02_ncu_general_summary.md"],
}


def dram_bytes(md_path):
    tot = 0.0
    for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        m = re.search(r"\| " + re.escape(name) + r" \| ([0-9.]+) \| (\w+) \|", open(md_path).read())
        if not m:
            raise ValueError(f"{name} not found in {md_path}")
        tot += float(m.group(1)) * UNIT[m.group(2)]
    return int(round(tot))


def build(commit):
    out = {}
    for key, files in SOURCES.items():
        out[key] = {"bytes": sum(dram_bytes(os.path.join(PROF, f)) for f in files), "captured_at": commit,
                    "source": [f"profiles/{f}" for f in files]}
    return out


if __name__ == "__main__":
    commit = sys.argv[1] if len(sys.argv) > 1 else subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, stdout=subprocess.PIPE, text=True).stdout.strip()
    d = build(commit)
    json.dump(d, open(os.path.join(PROF, "traffic.json"), "w"), indent=1)
    print(json.dumps(d, indent=1))
