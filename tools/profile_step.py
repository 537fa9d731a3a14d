"""Profiling driver: build the bench workload's engine and run a few eager (non-graph) steps so that ncu sees every
kernel launch by name.  Usage: ncu --profile-from-start off ... python tools/profile_step.py [--steps N_WARMUP]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "quantization-on-3d-object-detection_b200"))
import numpy as np
import torch

import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--graph", type=int, default=0)
    args = ap.parse_args()
    torch.cuda.set_device(0)
    pts = np.concatenate([np.concatenate([np.full((f.shape[0], 1), i, np.float32), f[:, 1:]], axis=1)
                          for i, f in enumerate(bench.make_batch(1000 + i, 1) for i in range(bench.BATCH))])
    eng, _ = bench.build_engine(torch.device("cuda", 0), pts.shape[0])
    eng.use_graph = bool(args.graph)
    eng.set_points(torch.from_numpy(pts))
    for _ in range(args.steps):                        # warm-up (outside the profiled range)
        eng.forward_points()
    torch.cuda.synchronize()
    # ncu --profile-from-start off: only the launches of this one eager step are profiled
    torch.cuda.cudart().cudaProfilerStart()
    eng.forward_points()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("counts", eng.counts(), "overflow", eng.overflowed(), "kernels/step", eng.kernels_per_forward)


if __name__ == "__main__":
    main()
