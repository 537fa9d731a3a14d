"""A/B of the two conv kernels on one bench configuration: per-layer CUDA-event times of the engine's conv launches with the
register-gather kernel (csrc/spconv_warp.cu) off / default / also for 64 x 64 layers, and the graph-replay step time of each.

  python tools/warp_ab.py [--config 2] [--modes 0,1,2]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "quantization-on-3d-object-detection_b200"))
import numpy as np
import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--modes", default="0,1,2")
    ap.add_argument("--quant", default=None, help="w_bits,act_bits,cw override, e.g. 8,8,0 for W8A8-pt")
    args = ap.parse_args()
    import bench
    bench.select(args.config)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    pts_np = np.concatenate([np.concatenate([np.full((f.shape[0], 1), i, np.float32), f[:, 1:]], axis=1)
                             for i, f in enumerate(bench.make_batch(1000 + fr, 1) for fr in range(bench.BATCH))])
    host = torch.from_numpy(pts_np).pin_memory()
    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    flush = lambda: flush_buf.zero_()
    res = {}
    for mode in args.modes.split(","):
        os.environ["QL_SPCONV_WARP"] = mode
        if args.quant:
            wb, ab, cw = [int(v) for v in args.quant.split(",")]
            eng, bb = bench.build_engine(dev, pts_np.shape[0], wb, ab, bool(cw))
        else:
            eng, bb = bench.build_engine(dev, pts_np.shape[0])
        eng.set_points(host)
        for _ in range(3):
            eng.forward_points()
        torch.cuda.synchronize()
        if eng.frame_cap_exceeded():
            eng.use_hash_voxelizer()
            for _ in range(3):
                eng.forward_points()
            torch.cuda.synchronize()
        ms = sorted(bench.timed_steps(eng, 20, flush, 1, None))
        times = eng.profile_ops(from_points=True, iters=5, flush=flush)
        conv = {k.split(":", 1)[1]: v for k, v in times.items() if k.startswith("conv:")}
        res[mode] = dict(step_ms=ms[len(ms) // 2], conv_ms=sum(conv.values()), layers=conv)
        print(f"mode {mode}: step {res[mode]['step_ms']:.4f} ms  conv {res[mode]['conv_ms']:.4f} ms", flush=True)
        del eng, bb
        torch.cuda.empty_cache()
    names = list(next(iter(res.values()))["layers"])
    print("layer".ljust(20) + "".join(f"mode {m}".rjust(10) for m in res))
    for n in names:
        print(n.ljust(20) + "".join(f"{res[m]['layers'][n] * 1e3:10.1f}" for m in res))
    print(json.dumps(res))


if __name__ == "__main__":
    main()
