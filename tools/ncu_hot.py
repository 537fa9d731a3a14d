"""Hot SASS instructions (warp-stall samples) of one launch of an .ncu-rep.  Usage: python tools/ncu_hot.py rep launch_index [top]"""
import csv, subprocess, sys
rep, kid = sys.argv[1], int(sys.argv[2])
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", ], capture_output=True, text=True).stdout
# split per kernel
blocks, cur = [], []
for line in out.splitlines():
    if line.startswith('"Kernel Name"'):
        if cur: blocks.append(cur)
        cur = [line]
    else:
        cur.append(line)
if cur: blocks.append(cur)
b = blocks[kid]
print(b[0][:120])
rows = list(csv.reader(b[1:]))
h = rows[0]
si = h.index("# Samples"); src = h.index("Source")
stalls = [(i, n) for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
data = rows[1:]
tot = sum(int(r[si]) for r in data)
print("total samples", tot, "instructions", len(data))
order = sorted(range(len(data)), key=lambda i: -int(data[i][si]))[:top]
for i in sorted(order):
    r = data[i]
    st = sorted(((int(r[j]), n) for j, n in stalls), reverse=True)[:2]
    print(f"{i:5d} {int(r[si]):6d} {100*int(r[si])/tot:5.1f}%  {r[src].strip()[:70]:70s} {st}")
