"""Experiment: the batch cut into S independent sub-batches, each with its own BackboneEngine (own buffers, own CUDA graph), replayed
concurrently on S streams -- so that one sub-batch's conv launches fill the SMs while another's drain / start up (a third of the conv time
of a step is per-launch latency: DESIGN.md 5e).  Prints the step time of the one-engine schedule and of the split schedules.

  python tools/dual_ab.py [--config 2] [--splits 1,2,4]
"""
import argparse
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
    ap.add_argument("--splits", default="1,2,4")
    ap.add_argument("--steps", type=int, default=30)
    args = ap.parse_args()
    import bench
    import qlidar
    bench.select(args.config)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    frames = [bench.make_batch(1000 + fr, 1) for fr in range(bench.BATCH)]
    bb, c = bench.build_backbone(dev)
    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    main_s = torch.cuda.current_stream()
    for S in [int(v) for v in args.splits.split(",")]:
        if bench.BATCH % S:
            continue
        per = bench.BATCH // S
        engs, streams = [], []
        for s in range(S):
            pts = np.concatenate([np.concatenate([np.full((f.shape[0], 1), i, np.float32), f[:, 1:]], axis=1)
                                  for i, f in enumerate(frames[s * per:(s + 1) * per])])
            cap = per * c["max_voxels"]
            eng = qlidar.BackboneEngine(bb, per, cap, max_points=pts.shape[0], pc_range=c["pc_range"], voxel_size=c["voxel_size"],
                                        max_pts_per_voxel=c["max_pts"], use_graph=True, device=dev, max_voxels_per_frame=c["max_voxels"],
                                        stage_caps=bench.stage_caps_for(cap), sorted_voxelizer=True)
            eng.set_points(torch.from_numpy(pts).pin_memory())
            for _ in range(3):
                eng.forward_points()
            torch.cuda.synchronize()
            if eng.frame_cap_exceeded():
                eng.use_hash_voxelizer()
                for _ in range(3):
                    eng.forward_points()
                torch.cuda.synchronize()
            assert not eng.overflowed()
            engs.append(eng)
            streams.append(torch.cuda.Stream(device=dev))

        def step():
            for e, st in zip(engs, streams):
                st.wait_stream(main_s)
                with torch.cuda.stream(st):
                    e.forward_points()
            for st in streams:
                main_s.wait_stream(st)

        for _ in range(5):
            step()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for a, b in ev:
            flush_buf.zero_()
            a.record()
            step()
            b.record()
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in ev)
        med = ms[len(ms) // 2]
        print(f"splits {S}: {per} frame(s) per engine, step {med:.4f} ms (p10 {ms[len(ms) // 10]:.4f}, p90 {ms[9 * len(ms) // 10]:.4f}) "
              f"= {bench.BATCH / med * 1e3:.1f} frames/s", flush=True)
        del engs
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
