"""BASELINE config 5 -- SubMConv3d layer sweep: C_in = C_out in {16..256}, kernel 3^3, N active voxels 10 k .. 2 M on the G2
surface-sheet generator, modes W8A16 / fp16 (kind::f16: fp16 rows x int8-code or fp16 weights -- same kernel,
same time), W8A8-pt (kind::i8, int8 codes, INT32 accumulate; reported with and without the activation quantiser in front).

Per point: kernel time (CUDA events on the launching stream, L2 flushed between iterations), algorithmic GB/s and TFLOP/s /
TOPS (SURVEY.md 8d formulas) and their fractions of the measured HBM peak and of 2x the measured bf16 peak for INT8
(MEASURED_PEAKS.json holds no INT8 number; nominal dense INT8 is 2x bf16).

  python tools/layer_sweep.py [--iters 5] [--md profiles/r01_layer_sweep.md]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "quantization-on-3d-object-detection_b200"))
import numpy as np
import torch


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d["bf16_tflops"]), "measured"
    return 6650.0, 1590.0, "fallback"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--md", default=None)
    ap.add_argument("--sizes", default="10000,30000,100000,300000,1000000,2000000")
    ap.add_argument("--channels", default="16,32,64,128,256")
    args = ap.parse_args()
    from qlidar import ops, synth
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    hbm, bf16, src = peaks()
    flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    rng = np.random.default_rng(2000)
    lines = ["| N | pairs/row | live offsets per tile (raster / key-sorted / grouped) | C | mode | conv us (raster order) | key-sorted us | grouped us | +quantise us | GB/s alg (best) | frac HBM | TFLOP/s (TOPS) alg (best) | frac tensor |",
             "|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    rows = []
    for N in [int(v) for v in args.sizes.split(",")]:
        S = int(round(N ** 0.5))
        coords_np = synth.synth_surface_sheet(S, seed=2000)
        n = coords_np.shape[0]
        coords = torch.from_numpy(coords_np).to(dev)
        grid = (1, 41, S, S)
        table = ops.hash_build(coords, None, grid)
        nbr, kmask = ops.rulebook_subm(coords, None, grid, 3, table, with_mask=True)
        pairs = int((nbr >= 0).sum().item())
        # the same sites renumbered by key (what the engine does for stage 1) and, on top, the grouped rulebook
        ws = torch.zeros(ops.rulebook_strided_workspace_bytes(grid, 1, 1, 0), dtype=torch.uint8, device=dev)
        oc, n_out, _, _ = ops.renumber_by_key(coords, None, grid, ws)
        index = ops.rulebook_strided_index(grid, 1, 1, 0, ws)
        nbr_s, kmask_s = ops.rulebook_subm_ranked(oc, n_out, grid, 3, index)
        nbr_g, kmask_g, perm_g = ops.rulebook_subm_ranked_grouped(oc, n_out, grid, 3, index)
        tiles = ops.num_tiles(n)
        live = lambda km: float(sum(bin(int(v) & 0xFFFFFFFF).count("1") for v in km[:tiles].cpu().numpy().ravel())) / tiles
        live_txt = f"{live(kmask):.1f} / {live(kmask_s):.1f} / {live(kmask_g):.1f}"
        for C in [int(v) for v in args.channels.split(",")]:
            x = rng.normal(size=(n, C)).astype(np.float32)
            x[::100, 1 % C] *= 20
            x[::100, 7 % C] *= 20                                    # 1 % x20 outliers in two fixed channels
            xf = torch.from_numpy(x).to(dev).half()
            w = torch.from_numpy(rng.integers(-127, 128, size=(C, 27, C)).astype(np.float32))
            scale = torch.full((C,), 1e-3, device=dev)
            shift = torch.zeros(C, device=dev)
            for mode in ("W8A16/fp16", "W8A8-pt"):
                i8 = mode == "W8A8-pt"
                packed = ops.pack_weights(w.to(torch.int8) if i8 else w.half()).to(dev)
                out = torch.empty((n, C), dtype=torch.float16, device=dev)
                if i8:
                    absmax = ops.absmax_cols(xf)
                    q, act_scale = ops.quantize_rows(xf, absmax, ops.QL_Q_CODES_PER_TENSOR)
                feats = q if i8 else xf
                t_conv, t_q = [], []
                for it in range(args.iters + 2):
                    flush_buf.zero_()
                    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                    e[0].record()
                    if i8:
                        absmax.zero_()
                        ops.absmax_cols(xf, absmax=absmax)
                        ops.quantize_rows(xf, absmax, ops.QL_Q_CODES_PER_TENSOR, out=q, act_scale=act_scale)
                    e[1].record()
                    ops.spconv_mma(feats, nbr, n, None, C, packed, scale, shift, act_scale=act_scale if i8 else None, relu=True, out=out, kmask=kmask)
                    e[2].record()
                    torch.cuda.synchronize()
                    if it >= 2:
                        t_q.append(e[0].elapsed_time(e[1]) * 1e3)
                        t_conv.append(e[1].elapsed_time(e[2]) * 1e3)
                tc, tq = float(np.median(t_conv)), float(np.median(t_q))
                # tilings: key-sorted rows, and key-sorted + grouped by line key (bit-identical results)
                alt = []
                for nb_, km_, pm_ in ((nbr_s, kmask_s, None), (nbr_g, kmask_g, perm_g)):
                    ts = []
                    for it in range(args.iters + 1):
                        flush_buf.zero_()
                        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        a.record()
                        ops.spconv_mma(feats, nb_, n, None, C, packed, scale, shift, act_scale=act_scale if i8 else None, relu=True, out=out, kmask=km_, row_perm=pm_)
                        b.record()
                        torch.cuda.synchronize()
                        if it >= 1:
                            ts.append(a.elapsed_time(b) * 1e3)
                    alt.append(float(np.median(ts)))
                b_act = 1 if i8 else 2
                bytes_alg = n * C * b_act + n * C * 2 + 27 * C * C * b_act + 4 * pairs
                flops = 2.0 * pairs * C * C
                tb = min(tc, alt[0], alt[1])
                gbs = bytes_alg / (tb * 1e-6) / 1e9
                tf = flops / (tb * 1e-6) / 1e12
                tpk = 2 * bf16 if i8 else bf16
                rows.append(dict(N=n, C=C, mode=mode, conv_us=tc, sorted_us=alt[0], grouped_us=alt[1], quant_us=tq, gbs=gbs, tflops=tf))
                lines.append(f"| {n} | {pairs / n:.1f} | {live_txt} | {C} | {mode} | {tc:.1f} | {alt[0]:.1f} | {alt[1]:.1f} | {tq:.1f} | {gbs:.0f} | {gbs / hbm:.3f} | {tf:.1f} | {tf / tpk:.3f} |")
                print(lines[-1], flush=True)
    head = (f"# SubMConv3d layer sweep (BASELINE config 5), B200\n\n`python tools/layer_sweep.py --iters {args.iters}`; peaks ({src}): HBM {hbm} GB/s, "
            f"bf16 {bf16} TFLOP/s (INT8 fraction against 2x that); kernel = `k_spconv_ts`, L2 flushed between iterations, medians.\n"
            "Three tilings of the same sites (bit-identical results): the generator's raster order (hash rulebook), rows renumbered by key (`ql_renumber_by_key`, "
            "what the engine does for stage 1) and key-sorted + grouped by line key (`ql_rulebook_subm_ranked_grouped`); GB/s and TFLOP/s use the best of the three.\n"
            "GB/s and TFLOP/s are ALGORITHMIC (SURVEY.md 8d): every distinct input row once, every output row once, weights once, 4 B per pair; 2*pairs*C*C flops.\n\n")
    if args.md:
        open(args.md, "w").write(head + "\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
