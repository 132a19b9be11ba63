"""Per-source-line static SASS instruction counts from `nvdisasm -g -c` output (developer tool)."""
import collections
import re
import sys

cur = None
per = collections.Counter()
ops = collections.defaultdict(collections.Counter)
total = 0
for line in open(sys.argv[1]):
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and cur:
        per[cur] += 1
        ops[cur][m.group(2).split('.')[0]] += 1
        total += 1
print("total static instructions", total)
want = sys.argv[2] if len(sys.argv) > 2 else None
for (f, l), n in sorted(per.items()):
    if want and f != want:
        continue
    print(f"{f}:{l:4d} {n:5d}  " + " ".join(f"{k}:{v}" for k, v in ops[(f, l)].most_common(6)))
