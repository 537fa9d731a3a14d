"""Summarise an .ncu-rep (read here, no GPU): one line per profiled launch with the counters DESIGN.md / bench.py cite.
Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--md]"""
import csv
import subprocess
import sys

WANT = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("gpu__time_duration.sum", "us"),
        ("dram__bytes_read.sum", "dram_rd_MB"), ("dram__bytes_write.sum", "dram_wr_MB"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1_wavefront%"),
        ("l1tex__t_sector_hit_rate.pct", "l1_hit%"), ("lts__t_sector_hit_rate.pct", "l2_hit%"),
        ("l1tex__m_xbar2l1tex_read_bytes.sum", "xbar_rd_MB"),
        ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor%"),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu_issue%"),
        ("smsp__inst_executed.sum", "inst"), ("launch__registers_per_thread", "regs"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%")]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    def find(k):                                       # some sections prefix their metrics ("TPC.TriageCompute.sm__pipe_tensor...")
        if k in h:
            return h.index(k)
        m = [i for i, x in enumerate(h) if x.endswith("." + k)]
        return m[0] if m else None
    idx = [(find(k), n, units[find(k)]) for k, n in WANT if find(k) is not None]
    md = "--md" in sys.argv
    names = [n for _, n, _ in idx]
    print(("| " + " | ".join(names) + " |") if md else "\t".join(names))
    if md:
        print("|" + "---|" * len(names))
    for r in rows[2:]:
        vals = []
        for i, n, u in idx:
            v = r[i]
            if n == "kernel":
                v = v.replace("void <unnamed>::", "").split("(")[0]
            elif n == "us":
                v = f"{float(v) / (1000.0 if u == 'ns' else 1.0):.1f}"
            elif n.endswith("_MB"):
                f = float(v)
                f = f / 1e6 if u == "byte" else (f / 1e3 if u == "Kbyte" else (f * 1e3 if u == "Gbyte" else f))
                v = f"{f:.1f}"
            elif n not in ("grid",):
                try:
                    v = f"{float(v):.1f}"
                except ValueError:
                    pass
            vals.append(v)
        print(("| " + " | ".join(vals) + " |") if md else "\t".join(vals))


if __name__ == "__main__":
    main()
