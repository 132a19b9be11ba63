"""Basic blocks of a kernel in `nvdisasm -c` output: size and opcode histogram of the largest ones (developer tool)."""
import collections
import re
import sys

blocks = []
cur = {"label": "entry", "ops": []}
func = None
for line in open(sys.argv[1]):
    m = re.match(r'\s*\.text\.(\S+):', line)
    if m:
        func = m.group(1)
    m = re.match(r'(\.L_x_\d+):', line.strip())
    if m:
        blocks.append(cur)
        cur = {"label": m.group(1), "ops": [], "func": func}
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m:
        cur.setdefault("func", func)
        cur["ops"].append(m.group(2))
        if m.group(2).split('.')[0] in ("BRA", "EXIT", "RET", "BRX", "CALL") and not m.group(1):
            blocks.append(cur)
            cur = {"label": cur["label"] + "+", "ops": [], "func": func}
blocks.append(cur)
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8
filt = sys.argv[3] if len(sys.argv) > 3 else ""
big = sorted([b for b in blocks if filt in (b.get("func") or "")], key=lambda b: -len(b["ops"]))[:n]
for b in big:
    h = collections.Counter(o.split('.')[0] for o in b["ops"])
    print(f'{b.get("func","")[:40]} {b["label"]:12s} {len(b["ops"]):5d}  ' + " ".join(f"{k}:{v}" for k, v in h.most_common(30)))
