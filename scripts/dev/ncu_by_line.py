"""Executed warp instructions per CUDA source line: joins an ncu source-page CSV (per SASS address) with `nvdisasm -g`
line info of the same kernel (developer tool).

    ncu -i X.ncu-rep --page source --csv > /tmp/x.csv
    cuobjdump -xelf all lib.so; nvdisasm -g -c kob_api.sm_100a.cubin > /tmp/lib.sass
    python scripts/dev/ncu_by_line.py /tmp/x.csv /tmp/lib.sass '_ZN3kob13kob_step_fastILi6ELi1ELb0' [rows]
"""
import collections
import csv
import re
import sys

csv_path, sass_path, mangled = sys.argv[1:4]
rows_n = float(sys.argv[4]) if len(sys.argv) > 4 else 1.0
# --- nvdisasm: offset -> (file, line) for the chosen function
line_of = {}
cur = None
infunc = False
for ln in open(sass_path):
    m = re.match(r'\s*\.text\.(\S+):', ln)
    if m:
        infunc = m.group(1).startswith(mangled)
        continue
    if ln.startswith('\t.section') or ln.startswith('.section'):
        infunc = False
    if not infunc:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]+)\*/\s+', ln)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur
# --- ncu csv
rows = list(csv.reader(open(csv_path)))
for i, r in enumerate(rows):
    if r and r[0] == "Address":
        hdr, start = r, i + 1
        break
ie, ia, isrc = hdr.index("Instructions Executed"), hdr.index("Address"), hdr.index("Source")
base = int(rows[start][ia], 16)
per = collections.Counter()
ops = collections.defaultdict(collections.Counter)
tot = 0
for r in rows[start:]:
    try:
        n = int(r[ie])
    except Exception:
        continue
    off = int(r[ia], 16) - base
    key = line_of.get(off, ("?", 0))
    per[key] += n
    s = re.sub(r'^@!?U?P\d+\s+', '', r[isrc].strip())
    ops[key][s.split()[0].split('.')[0]] += n
    tot += n
print(f"total {tot}  per row {tot / rows_n:.1f}")
for (f, l), n in sorted(per.items()):
    if n / rows_n >= 0.3:
        print(f"{f}:{l:4d} {n / rows_n:7.1f}  " + " ".join(f"{k}:{v / rows_n:.1f}" for k, v in ops[(f, l)].most_common(5)))
