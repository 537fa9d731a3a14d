#!/usr/bin/env python
"""bench.py -- CenterPoint 3-D-backbone frames/s on B200 (BASELINE.json metric), one JSON line on stdout.

Workload (BASELINE.json configs[1]): CenterPoint Waymo VoxelResBackBone8x, W8A16 "progressive" quantisation
(QConvNd(w_bits=8, act_bits=16, cw=True) on every backbone conv except conv_input.0, as quant_centerpoint.quant with
sq=True), batch 4 synthetic ~146k-voxel frames per GPU.  One step = voxelize+meanVFE -> 9 rulebooks -> 21 fused convs ->
BEV densify for one batch.  N GPUs = N independent frame shards (rank r takes frames r::N, pcdet/datasets/__init__.py:45-49),
no collective on the hot path; NCCL only gathers the per-rank stage counts ("detections" placeholder) after timing.

  python bench.py [--gpus N] [--steps K] [--warmup W]              our arm (CUDA, C ABI)
  python bench.py --impl reference [--steps K] [--warmup W]         the reference's CPU path (oracle port) on the host cores
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "quantization-on-3d-object-detection_b200"))

import numpy as np
import torch

METRIC = "centerpoint_3d_backbone_frames_per_sec"
BATCH = 4
W_BITS, ACT_BITS, CW = 8, 16, True
NO_LIST = ["conv_input.0"]                     # quant/quant_centerpoint.py:24-26 (backbone_no_list, module-relative path)
WORKLOAD = "centerpoint_waymo_voxelresbackbone8x_w8a16_batch4_synthetic_146k_voxel_frames"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def make_batch(first_seed: int, batch: int = BATCH) -> np.ndarray:
    from qlidar import synth
    return synth.synth_batch("waymo", batch, first_seed=first_seed)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=float(d["hbm_gbs"]), bf16=float(d["bf16_tflops"]), bf16_sus=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sus=1400.0, src="fallback (B200_PROFILING.md)")


# tcgen05.mma kind::i8 ceiling of this pool's B200s (M = 128, N >= 128, A in TMEM, all 148 SMs issuing back to back):
# tools/microbench/mma_peak.cu, output committed as profiles/r01_mma_peak_microbench.txt (kind::f16: 2233 TFLOP/s)
I8_MMA_PEAK_TOPS = 4596.0
# experiment switch: extra BackboneEngine keyword arguments as JSON, e.g. QL_ENGINE_KW='{"group_rows": false}' (default: none)
ENGINE_KW = json.loads(os.environ.get("QL_ENGINE_KW", "{}"))


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md's clocks line).  NVML in-process
    (a 2 ms polling thread: the timed region is only ~50-100 ms, too short for an `nvidia-smi -lms` child to start up --
    an 8-GPU run got zero samples that way); `nvidia-smi` is the fallback when pynvml is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []
        self.nvml, self.h, self.samples, self._stop = None, None, [], threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = gpu_index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if gpu_index < len(ids) and ids[gpu_index].isdigit():
                    phys = int(ids[gpu_index])
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _poll(self):
        n = self.nvml
        get_reasons = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(n, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop.is_set():
            try:
                self.samples.append((float(n.nvmlDeviceGetClockInfo(self.h, n.NVML_CLOCK_SM)), int(get_reasons(self.h))))
            except Exception:
                break
            time.sleep(0.002)

    def start(self):
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self._stop.set()
            self.thread.join(timeout=1.0)
            n = self.nvml
            bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
            reasons = sorted(k for k, b in bits.items() if any(r & b for _, r in self.samples))
            sm = [c for c, _ in self.samples]
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.smax, "reasons": reasons,
                    "samples": len(sm), "source": "nvml"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi"}


# ----------------------------------------------------------------------------------------------------- our arm
def build_engine(device, max_points, w_bits=None, act_bits=None, cw=None):
    import qlidar
    w_bits = W_BITS if w_bits is None else w_bits
    act_bits = ACT_BITS if act_bits is None else act_bits
    cw = CW if cw is None else cw
    from qlidar import synth
    c = synth.CONFIGS["waymo"]
    r = np.asarray(c["pc_range"], dtype=np.float64)
    grid = np.round((r[3:6] - r[0:3]) / np.asarray(c["voxel_size"], dtype=np.float64)).astype(np.int64)
    torch.manual_seed(4)                                      # the reference's seed (quant_centerpoint.py:174)
    bb = qlidar.VoxelResBackBone8x({}, c["nfeat"], grid)
    with torch.no_grad():                                     # random-init BN statistics per SURVEY.md 8d
        for m in bb.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.weight.uniform_(0.5, 1.5); m.bias.normal_(0, 0.1); m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.5, 1.5)
    bb = bb.to(device).eval()
    qlidar.q_conv3d(bb, {}, "", w_bits, act_bits, cw, (qlidar.SubMConv3d, qlidar.SparseConv3d), NO_LIST)
    # every frame is capped at MAX_NUMBER_OF_VOXELS = 150 000 voxels in first-touch order like the reference's per-frame CPU
    # voxeliser (waymo_dataset.yaml:79-84; the synthetic frames hold 98 k - 160 k voxels depending on the seed), so the batch
    # capacity BATCH * 150 000 can never overflow
    cap = BATCH * c["max_voxels"]
    eng = qlidar.BackboneEngine(bb, BATCH, cap, max_points=max_points, pc_range=c["pc_range"], voxel_size=c["voxel_size"],
                                max_pts_per_voxel=c["max_pts"], use_graph=True, device=device, max_voxels_per_frame=c["max_voxels"],
                                stage_caps=[cap, int(1.25 * cap), int(0.75 * cap), int(0.5 * cap), int(0.5 * cap)], **ENGINE_KW)
    return eng, bb


def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from qlidar import ops

    # frames r::N of the job's N*BATCH frames (pcdet/datasets/__init__.py:40-50 via qlidar.shard): frame f has seed 1000 + f
    from qlidar import shard
    my_frames = shard.frames_for_rank(world * BATCH, rank, world)
    pts_np = np.concatenate([np.concatenate([np.full((f.shape[0], 1), i, np.float32), f[:, 1:]], axis=1)
                             for i, f in enumerate(make_batch(1000 + fr, 1) for fr in my_frames)])
    P = pts_np.shape[0]
    host_pts = [torch.from_numpy(pts_np).pin_memory(), torch.from_numpy(pts_np.copy()).pin_memory()]
    eng, _ = build_engine(dev, P)
    eng.set_points(host_pts[0])
    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def flush():
        flush_buf.zero_()

    # ---- warm-up (graph capture happens on the first replay) ----
    for _ in range(max(args.warmup, 3)):
        eng.forward_points()
    torch.cuda.synchronize()
    if eng.overflowed():
        raise SystemExit(f"bench.py: a stage capacity overflowed (rank {rank}, (kept, found) per stage {[st.n_dev.tolist() for st in eng.stages]}); raise stage_caps")
    counts = eng.counts()
    kernels_per_step = eng.kernels_per_forward

    # ---- timed region A: device-resident inputs, K steps, L2 flushed between steps, CUDA events per step ----
    sampler = ClockSampler(local)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        flush()
        ev[i][0].record()
        eng.forward_points()
        ev[i][1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.stop()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = float(sum(step_ms))
    t = torch.tensor([total_ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * BATCH * args.steps / (total_ms_max / 1e3)

    # ---- timed region B (e2e): pinned host points -> H2D -> step -> D2H of the stage counts, double buffered ----
    copy_stream = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream()
    stage_in = [torch.empty((P, pts_np.shape[1]), dtype=torch.float32, device=dev) for _ in range(2)]
    counts_dev = torch.zeros((len(eng.stages), 2), dtype=torch.int32, device=dev)
    counts_host = torch.zeros((len(eng.stages), 2), dtype=torch.int32).pin_memory()
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    e2e_steps = args.steps

    def h2d(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i & 1])
            stage_in[i & 1].copy_(host_pts[i & 1], non_blocking=True)
            copied[i & 1].record(copy_stream)

    for e in consumed:
        e.record(main)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_start.record(main)
    h2d(0)
    for i in range(e2e_steps):
        if i + 1 < e2e_steps:
            h2d(i + 1)                                         # overlaps the previous step's compute
        main.wait_event(copied[i & 1])
        eng.points[:P].copy_(stage_in[i & 1], non_blocking=True)   # device-side hand-off into the graph's static input
        consumed[i & 1].record(main)
        eng.forward_points()
        torch.stack([st.n_dev for st in eng.stages], out=counts_dev)
        counts_host.copy_(counts_dev, non_blocking=True)       # the step's host-visible result
    e_end.record(main)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e2e_ms = torch.tensor([e_start.elapsed_time(e_end)], device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * BATCH * e2e_steps / (float(e2e_ms.item()) / 1e3)
    assert counts_host[:, 0].tolist() == counts, "e2e path disagrees with the device-resident path"

    # ---- NCCL only gathers detections (here: per-rank stage counts) after the timed regions ----
    gathered = None
    if world > 1:
        # one fixed-shape block per frame of this rank (here the batch's stage counts stand in for padded detections),
        # merged back into dataset order: common_utils.py:229-250 without the pickle files
        mine = counts_dev.flatten().to(torch.float32).unsqueeze(0).repeat(BATCH, 1).contiguous()
        gathered = shard.gather_frame_results(mine, world * BATCH)[:world].to(torch.int64).cpu().tolist()   # frame r belongs to rank r

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-kernel accounting (rank 0): eager pass, CUDA events around every op on the launching stream ----
    pk = peaks()
    times = eng.profile_ops(from_points=True, iters=5, flush=flush)
    acct = {a["name"]: a for a in eng.layer_accounting()}
    conv_ms = {k.split(":", 1)[1]: v for k, v in times.items() if k.startswith("conv:")}
    conv_bytes = sum(acct[n]["bytes_alg"] for n in conv_ms)
    conv_flops = sum(acct[n]["flops_alg"] for n in conv_ms)
    conv_t = sum(conv_ms.values()) / 1e3
    groups = {}
    for k, v in times.items():
        g = k.split(":")[0]
        groups[g] = groups.get(g, 0.0) + v
    eager_total = sum(times.values())
    n0, P_in = counts[0], P
    F = eng.nfeat
    vox_bytes = P_in * (F + 1) * 4 + n0 * F * 4 + n0 * 16
    last = eng.stages[-1]
    bev_bytes = counts[-1] * 128 * 2 + int(np.prod(eng.spatial_features.shape)) * 2
    rb_bytes = 0
    for L in eng.layers:
        pass
    stage_gbs = {
        "voxelize_mean": vox_bytes / (times["voxelize_mean"] / 1e3) / 1e9,
        "bev_densify": bev_bytes / (times["bev_densify"] / 1e3) / 1e9,
        "spconv_mma_all_layers": conv_bytes / conv_t / 1e9,
    }
    per_layer = [dict(name=n, ms=round(conv_ms[n], 4), gbs=round(acct[n]["bytes_alg"] / (conv_ms[n] / 1e3) / 1e9, 1),
                      tflops=round(acct[n]["flops_alg"] / (conv_ms[n] / 1e3) / 1e12, 2), cin=acct[n]["cin"], cout=acct[n]["cout"],
                      n_out=acct[n]["n_out"], pairs=acct[n]["pairs"]) for n in conv_ms]
    # measured DRAM traffic of the same 20 launches from the committed ncu capture (profiles/, not measured in this run)
    traffic, traffic_src = None, None
    tp = os.path.join(ROOT, "profiles", "r01_conv_traffic.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        traffic, traffic_src = int(tj["traffic_bytes_per_step"]), "profiles/r01_conv_traffic.json (ncu --set full, dram__bytes_read+write over the 20 launches)"
    roofline = {"kernel": "k_spconv_ts<f16> (the 20 tcgen05 sparse-conv launches of one step, aggregated)", "bound": "hbm",
                "achieved": round(conv_bytes / conv_t / 1e9, 1), "peak": pk["hbm"], "unit": "GB/s",
                "frac": round(conv_bytes / conv_t / 1e9 / pk["hbm"], 4), "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": pk["src"],
                "alg_bytes_per_step": conv_bytes, "kernel_ms_per_step": round(conv_t * 1e3, 3),
                "share_of_step": round(conv_t * 1e3 / eager_total, 3),
                "tensor_tflops_alg": round(conv_flops / conv_t / 1e12, 2), "tensor_frac_of_bf16_peak": round(conv_flops / conv_t / 1e12 / pk["bf16"], 4)}

    # ---- INT8 leg (BASELINE metric: "INT8 sparse-conv TOPS"): the same backbone as W8A8 per-tensor (QConvNd(8, 8, cw=False):
    #      int8 codes x int8 codes -> INT32 on tcgen05 kind::i8, dynamic abs-max fused into the producing epilogue) ----
    int8_leg = None
    try:
        import qlidar
        del eng
        torch.cuda.empty_cache()

        def time_engine(e8):
            e8.set_points(host_pts[0])
            for _ in range(3):
                e8.forward_points()
            torch.cuda.synchronize()
            ev8 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
            for i in range(args.steps):
                flush()
                ev8[i][0].record()
                e8.forward_points()
                ev8[i][1].record()
            torch.cuda.synchronize()
            ms8 = float(np.median([a.elapsed_time(b) for a, b in ev8]))
            t8 = e8.profile_ops(from_points=True, iters=3, flush=flush)
            acct8 = {a["name"]: a for a in e8.layer_accounting()}
            c8 = {k.split(":", 1)[1]: v for k, v in t8.items() if k.startswith("conv:")}
            i8_names = [n for n in c8 if acct8[n]["kind"] == "i8"]
            ops8 = sum(acct8[n]["flops_alg"] for n in i8_names)
            tc8 = sum(c8[n] for n in i8_names) / 1e3
            q8 = sum(v for k, v in t8.items() if k.startswith("quantize:")) / 1e3
            return {"frames_per_sec": round(BATCH / (ms8 / 1e3), 2), "ms_per_step": round(ms8, 4), "conv_ms_per_step": round(tc8 * 1e3, 4),
                    "quantize_ms_per_step": round(q8 * 1e3, 4), "tops_alg": round(ops8 / tc8 / 1e12, 2),
                    "frac_of_2x_bf16_peak": round(ops8 / tc8 / 1e12 / (2 * pk["bf16"]), 4),
                    "frac_of_i8_mma_peak": round(ops8 / tc8 / 1e12 / I8_MMA_PEAK_TOPS, 4)}

        eng8, bb8 = build_engine(dev, P, 8, 8, False)
        dyn = time_engine(eng8)
        # static calibration exactly as the reference drivers do it (collect_stats -> compute_amax, quant/quantize.py:175-207),
        # one batch through the eager module path, then a second engine that consumes the frozen amax tables
        n0 = eng8.counts()[0]
        calib = {"voxel_features": eng8.vox_feats[:n0, :eng8.nfeat].clone(), "voxel_coords": eng8.stages[0].coords[:n0].float(), "batch_size": BATCH}
        del eng8
        torch.cuda.empty_cache()

        class _Pipe(torch.nn.Module):
            def __init__(self, m):
                super().__init__()
                self.backbone_3d = m

            def forward(self, bd):
                return self.backbone_3d(bd)

        qlidar.collect_stats(_Pipe(bb8), [calib], n_batches=0)
        qlidar.compute_amax(bb8, dev)
        c = __import__("qlidar").synth.CONFIGS["waymo"]
        cap = BATCH * c["max_voxels"]
        eng8s = qlidar.BackboneEngine(bb8, BATCH, cap, max_points=P, pc_range=c["pc_range"], voxel_size=c["voxel_size"],
                                      max_pts_per_voxel=c["max_pts"], use_graph=True, device=dev, max_voxels_per_frame=c["max_voxels"],
                                      stage_caps=[cap, int(1.25 * cap), int(0.75 * cap), int(0.5 * cap), int(0.5 * cap)], **ENGINE_KW)
        sta = time_engine(eng8s)
        sta["fused_requantised_layers"] = int(sum(L.fused_q for L in eng8s.layers))
        int8_leg = {"mode": "QConvNd(w_bits=8, act_bits=8, cw=False): W8A8 per-tensor, INT32 accumulate (tcgen05 kind::i8)",
                    "dynamic_amax": dyn, "static_calibration": sta,
                    "note": "INT8 peak is not in MEASURED_PEAKS.json; fractions against 2x the measured bf16 (cuBLAS) peak and against the "
                            "measured tcgen05 kind::i8 issue ceiling (4596 TOPS, profiles/r01_mma_peak_microbench.txt). "
                            "static = collect_stats/compute_amax on one batch, int8 codes written by the producing layer's epilogue"}
    except Exception as e:                                     # the headline line must not depend on the extra leg
        int8_leg = {"error": repr(e)[:200]}

    # ---- CPU baseline beside it: the oracle port on ONE frame of the same workload ----
    cpu = cpu_baseline(pts_np, sample_frames=3, warm=0)        # ~10 s of host work: 3 of the batch's 4 frames

    line = {
        "metric": METRIC, "value": round(value, 2), "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": round(total_ms_max / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 activations x int8-code weights, fp32 accumulate (W8A16)", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": BATCH, "points_per_step": int(P), "voxels_per_stage": counts,
                   "l2": "512 MiB flush between timed steps", "quant": "QConvNd(w_bits=8, act_bits=16, cw=True), conv_input.0 unquantised",
                   "parallelism": f"frame-sharded x{world} (no data-path collective)"},
        "e2e": {"value": round(e2e_value, 2), "unit": "frames/s", "h2d_bytes_per_step": int(pts_np.nbytes),
                "d2h_bytes_per_step": int(counts_host.numel() * 4), "ms_per_step": round(float(e2e_ms.item()) / e2e_steps, 4),
                "note": "pinned host points -> H2D (copy stream, double buffered) -> graph replay -> D2H stage counts"},
        "gpu_launches": int(kernels_per_step * args.steps),
        "kernels_per_step": int(kernels_per_step),
        "clocks": clocks,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "stage_ms_eager": {k: round(v, 4) for k, v in groups.items()},
        "stage_gbs": {k: round(v, 1) for k, v in stage_gbs.items()},
        "stage_frac_of_hbm_peak": {k: round(v / pk["hbm"], 4) for k, v in stage_gbs.items()},
        "conv_layers": per_layer,
        "int8": int8_leg,
        "step_ms_p10_p50_p90": [round(float(np.percentile(step_ms, q)), 4) for q in (10, 50, 90)],
    }
    if gathered is not None:
        line["gathered_stage_counts"] = gathered
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------- CPU arms
def cpu_baseline(pts_batch: np.ndarray, sample_frames: int = 1, warm: int = 0):
    """The reference's CPU path restated (oracle/): per-frame hard voxelisation + MeanVFE + VoxelResBackBone8x with the
    reference's fake-quant math (QConvNd, quant/quant.py:36-58) + HeightCompression, all host cores."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import qlidar_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    c = O.CONFIGS["waymo"]
    grid = O.grid_size_xyz(c["pc_range"], c["voxel_size"])
    prog = O.backbone_specs("VoxelResBackBone8x", c["nfeat"])
    params = O.init_params(prog)
    q = O.QuantCfg(mode="ref", w_bits=W_BITS, act_bits=ACT_BITS, cw=CW, no_list=tuple(NO_LIST), fast=True)
    frames = [pts_batch[pts_batch[:, 0] == b] for b in range(int(pts_batch[:, 0].max()) + 1)]

    def one(f):
        f = f.copy(); f[:, 0] = 0
        feats, coords, _ = O.voxelize_mean_batch(f, c["pc_range"], c["voxel_size"], c["max_pts"], c["max_voxels"])
        out, _ = O.backbone_forward(prog, params, torch.from_numpy(feats), coords, O.sparse_shape_zyx(grid), 1, q)
        O.height_compression(out.features, out.coords, out.spatial_shape, 1)
        return coords.shape[0]

    for i in range(warm):
        one(frames[i % len(frames)])
    t0 = time.perf_counter()
    nv = [one(frames[i % len(frames)]) for i in range(sample_frames)]
    dt = time.perf_counter() - t0
    model = "unknown"
    try:
        for l in open("/proc/cpuinfo"):
            if l.startswith("model name"):
                model = l.split(":", 1)[1].strip(); break
    except OSError:
        pass
    return {"value": round(sample_frames / dt, 5), "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{sample_frames} frame(s) of the same batch ({nv[0]} voxels), torch CPU with {cores} threads, {model}",
            "seconds": round(dt, 2)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return                                                # other ranks exit 0 without work
    budget_s = 150.0
    pts = make_batch(1000, BATCH)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import qlidar_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    c = O.CONFIGS["waymo"]
    grid = O.grid_size_xyz(c["pc_range"], c["voxel_size"])
    prog = O.backbone_specs("VoxelResBackBone8x", c["nfeat"])
    params = O.init_params(prog)
    q = O.QuantCfg(mode="ref", w_bits=W_BITS, act_bits=ACT_BITS, cw=CW, no_list=tuple(NO_LIST), fast=True)
    frames = []
    for b in range(BATCH):
        f = pts[pts[:, 0] == b].copy(); f[:, 0] = 0
        frames.append(f)

    def one(f):
        feats, coords, _ = O.voxelize_mean_batch(f, c["pc_range"], c["voxel_size"], c["max_pts"], c["max_voxels"])
        out, _ = O.backbone_forward(prog, params, torch.from_numpy(feats), coords, O.sparse_shape_zyx(grid), 1, q)
        O.height_compression(out.features, out.coords, out.spatial_shape, 1)
        return coords.shape[0]

    t_w = time.perf_counter()
    nv = one(frames[0])                                        # 1 warm-up frame (also sizes the budget)
    per = time.perf_counter() - t_w
    warm_done = 1
    while warm_done < min(args.warmup, 1 + int(0.2 * budget_s / per)):
        one(frames[warm_done % BATCH]); warm_done += 1
    steps = max(1, min(args.steps, int(budget_s / per)))
    t0 = time.perf_counter()
    for i in range(steps):
        one(frames[i % BATCH])                                 # one step = ONE frame of the batch (bounded sample)
    dt = time.perf_counter() - t0
    v = steps / dt
    base = {"value": round(v, 5), "unit": "frames/s", "cores": cores, "kind": "port",
            "sample": f"{steps} step(s) of 1 frame ({nv} voxels) each; requested {args.steps} steps, bounded to ~{int(budget_s)} s"}
    line = {"impl": "reference", "metric": METRIC, "value": round(v, 5), "unit": "frames/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
            "steps": steps, "warmup": warm_done, "ms_per_step": round(dt / steps * 1e3, 2), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32 fake-quant (reference math)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "reference's CPU path = oracle port (spconv / pytorch_quantization are not installable here)"},
            "cpu_baseline": base, "e2e": {"value": round(v, 5), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
