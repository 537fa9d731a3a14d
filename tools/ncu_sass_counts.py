"""Per-SASS-line execution counts and stall samples of one launch in an ncu report (source page).
Usage: python tools/ncu_sass_counts.py report.ncu-rep launch_index first_line last_line"""
import csv, subprocess, sys
rep, kid, a, b = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], []
for line in out.splitlines():
    if line.startswith('"Kernel Name"'):
        if cur: blocks.append(cur)
        cur = [line]
    else: cur.append(line)
if cur: blocks.append(cur)
rows = list(csv.reader(blocks[kid][1:])); h = rows[0]; data = rows[1:]
si = h.index("# Samples"); src = h.index("Source"); ex = h.index("Instructions Executed")
for i in range(a, b):
    r = data[i]
    print(i, r[ex], r[si], r[src][:110])
