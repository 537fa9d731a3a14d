"""Role wait-cycle trace of the conv kernel (needs the library built with profiles/r01_conv_trace.patch applied and
QL_SPCONV_TRACE=1): per conv launch of the bench workload, the share of its time each role's lead thread spends inside its
barrier waits -- the role that never waits is the one the others wait for."""
import ctypes as C
import os
import sys

os.environ["QL_SPCONV_TRACE"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "quantization-on-3d-object-detection_b200"))
import numpy as np
import torch
import bench
from qlidar import _lib

torch.cuda.set_device(0)
pts = np.concatenate([np.concatenate([np.full((f.shape[0], 1), i, np.float32), f[:, 1:]], axis=1)
                      for i, f in enumerate(bench.make_batch(1000 + i, 1) for i in range(bench.BATCH))])
eng, _ = bench.build_engine(torch.device("cuda", 0), pts.shape[0])
eng.use_graph = False
eng.overlap_rulebooks = False
eng.set_points(torch.from_numpy(pts))
eng.forward_points()
torch.cuda.synchronize()
lib = C.CDLL(_lib.LIB_PATH)
buf = (C.c_ulonglong * 16)()
lib.ql_debug_trace_read(buf, 1)
print("| layer | producer: wait rulebook / wait free slot | MMA thread: wait rulebook / wait accumulator / wait unit | epilogue: wait accumulator | loader: wait buffer / wait B slot |")
print("|---|---|---|---|---|")


def op(label, n, fn, *a, **kw):
    r = fn(*a, **kw)
    if label.startswith("conv:"):
        torch.cuda.synchronize()
        lib.ql_debug_trace_read(buf, 1)
        v = [float(x) for x in buf]
        pct = lambda a, b: f"{100.0 * a / max(b, 1.0):.0f} %"
        print(f"| {label[5:]} | {pct(v[1], v[0])} / {pct(v[2], v[0])} | {pct(v[4], v[3])} / {pct(v[5], v[3])} / {pct(v[6], v[3])} | {pct(v[8], v[7])} | "
              f"{pct(v[10], v[9])} / {pct(v[11], v[9])} |", flush=True)
    return r


eng._op = op
eng.forward_points()
torch.cuda.synchronize()
