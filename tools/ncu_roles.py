"""Per-role warp-stall sample totals for the warp-specialised conv kernel: the SASS is split at the role landmarks
(STTM = gather producers, LDTM = epilogue, UTC*MMA = MMA issuer, UBLKCP = loader).  Usage: python tools/ncu_roles.py rep launch_index"""
import csv, subprocess, sys, collections
rep, kid = sys.argv[1], int(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
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
rows = list(csv.reader(b[1:])); h = rows[0]; data = rows[1:]
si = h.index("# Samples"); src = h.index("Source"); ex = h.index("Instructions Executed")
stalls = [(i, n) for i, n in enumerate(h) if n.startswith("stall_") and "Not Issued" not in n]
marks = []
for i, r in enumerate(data):
    s = r[src]
    for key in ("STTM", "LDTM", "UTCHMMA", "UTCIMMA", "UTCQMMA", "UBLKCP", "BAR.SYNC", "EXIT"):
        if key in s: marks.append((i, key))
print("landmarks:", [(i, k) for i, k in marks][:60])
# windows given on the command line: a:b
for w in sys.argv[3:]:
    a, b2 = [int(x) for x in w.split(":")]
    tot = collections.Counter(); n = 0; inst = 0
    for r in data[a:b2]:
        n += int(r[si]); inst += int(r[ex])
        for j, nm in stalls: tot[nm] += int(r[j])
    print(w, "samples", n, "warp-inst", inst, tot.most_common(6))
