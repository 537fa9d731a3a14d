"""Tuning driver: per-layer conv times of the bench workload (eager, CUDA events per op) under a list of kernel-plan overrides.
Usage: python tools/conv_sweep.py "TEAMS=4" "TEAMS=3" "UNIT_SUBS=16,SLOTS=3" ...   (QL_SPCONV_<KEY>=<value>; "" = defaults)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "quantization-on-3d-object-detection_b200"))
import numpy as np
import torch

import bench

KEYS = ("UNIT_SUBS", "SLOTS", "TEAMS", "STREAM", "MODE", "CTAS", "ACC")
ROLES = {0: ("epilogue", ["wait_acc", "work"]), 1: ("producer", ["wait_desc", "gather", "wait_slot", "st+arrive"]),
         2: ("mma", ["wait_desc", "wait_acc", "wait_unit", "issue", "commit"]), 3: ("loader", ["wait_desc_free", "wait_slot", "work"])}


def print_trace(eng):
    """Role trace of the test-time build: share of each role's lifetime (lead thread, summed over the CTAs) per phase."""
    import ctypes
    from qlidar import _lib
    lib = _lib.lib()
    buf = (ctypes.c_ulonglong * (64 * 32))()
    lib.ql_debug_read_trace(buf)                      # clear
    eng._run_from_points()
    torch.cuda.synchronize()
    lib.ql_debug_read_trace(buf)
    convs = [L for L in eng.layers if L.kind != "stem"]
    for i, L in enumerate(convs):
        row = [buf[i * 32 + j] for j in range(32)]
        parts = []
        for r, (name, phases) in ROLES.items():
            tot = row[r * 8 + 7] or 1
            parts.append(name + " " + "/".join(f"{ph}={100.0 * row[r * 8 + k] / tot:.0f}%" for k, ph in enumerate(phases)))
        print(f"  {L.name:14s} " + " | ".join(parts), flush=True)



def main():
    torch.cuda.set_device(0)
    pts = np.concatenate([np.concatenate([np.full((f.shape[0], 1), i, np.float32), f[:, 1:]], axis=1)
                          for i, f in enumerate(bench.make_batch(1000 + i, 1) for i in range(bench.BATCH))])
    eng, _ = bench.build_engine(torch.device("cuda", 0), pts.shape[0])
    eng.set_points(torch.from_numpy(pts))
    eng.forward_points()
    torch.cuda.synchronize()
    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
    for cfg in (sys.argv[1:] or [""]):
        for k in KEYS:
            os.environ.pop("QL_SPCONV_" + k, None)
        abl, trace = 0, False
        for kv in filter(None, cfg.split(",")):
            k, v = kv.split("=")
            if k == "ABLATE":                     # needs QLIDAR_LIB=.../libqlidar_b200_ablate.so (build.py --ablate)
                abl = int(v)
            elif k == "TRACE":
                trace = bool(int(v))
            else:
                os.environ["QL_SPCONV_" + k] = v
        from qlidar import _lib
        if hasattr(_lib.lib(), "ql_debug_set_ablate"):
            _lib.lib().ql_debug_set_ablate(abl)
        elif abl:
            raise SystemExit("ABLATE needs the ablation build: QLIDAR_LIB=<...>/libqlidar_b200_ablate.so")
        t = eng.profile_ops(iters=4, flush=lambda: flush_buf.zero_())
        conv = {k.split(":", 1)[1]: v * 1e3 for k, v in t.items() if k.startswith("conv:")}
        tot = sum(conv.values())
        print(f"[{cfg or 'default'}] conv total {tot:.0f} us | " + " ".join(f"{k.replace('conv', 'c')}={v:.0f}" for k, v in conv.items()), flush=True)
        if trace:
            print_trace(eng)


if __name__ == "__main__":
    main()
