"""W8A8 per-tensor, whole backbone: the CUDA engine against the oracle's KERNEL-NUMERICS MIRROR (oracle/qlidar_oracle.py,
mirror_backbone_w8a8_pt + oracle/qloracle_c.c), layer by layer, with tolerance ZERO:

  * the int8 activation codes every quantised layer gathers,
  * its INT8 x INT8 -> INT32 accumulators,
  * the fp16 rows it stores,

through all 21 (VoxelResBackBone8x) / 12 (VoxelBackBone8x) layers, dynamic amax and static calibration (where 19 layers take
their codes from the producing conv's epilogue).  A plain fp32 restatement cannot be compared this tightly -- one
fp16-vs-fp32 ulp flips a round-half-even decision, 0.8 % of amax each, 20 layers deep; the mirror restates the inter-layer
arithmetic the way the device performs it (fp16 storage, one fp32 FMA for de-quantisation + BatchNorm, fp16 residual), so the
comparison is exact and the north star's "INT8xINT8->INT32 accumulators bit-exact" gate holds for the whole network, not
only for single layers.  The mirror itself is held to the reference-math oracle layer by layer (teacher forced) below."""
import numpy as np
import pytest
import torch

import qlidar_oracle as O
from test_gpu_backbone import make_frame, build, batch_dict

pytestmark = pytest.mark.gpu


def _sorted_frame(cfg, batch, **kw):
    pts, feats, coords, grid, c = make_frame(cfg, batch, **kw)
    order = np.argsort(O._lin(coords, O.sparse_shape_zyx(grid)), kind="stable")     # the engine's stage-1 order
    return feats[torch.from_numpy(order)].contiguous(), np.ascontiguousarray(coords[order]), grid, c


@pytest.mark.parametrize("static", [False, True])
@pytest.mark.parametrize("arch,cfg,nfeat", [("VoxelResBackBone8x", "waymo", 5), ("VoxelBackBone8x", "kitti", 4)])
def test_w8a8_pt_codes_accumulators_and_rows_bit_exact_through_all_layers(arch, cfg, nfeat, static):
    import qlidar
    from qlidar import ops
    batch = 2
    feats, coords, grid, c = _sorted_frame(cfg, batch)
    prog, P, bb = build(arch, nfeat, grid)
    no_list = ["conv_input.0"]
    qlidar.q_conv3d(bb, {}, "", 8, 8, False, (qlidar.SubMConv3d, qlidar.SparseConv3d), no_list)
    amax = None
    if static:
        class Pipe(torch.nn.Module):
            def __init__(self):
                super().__init__()
                self.backbone_3d = bb

            def forward(self, bd):
                return self.backbone_3d(bd)

        qlidar.collect_stats(Pipe(), [batch_dict(feats, coords, batch)], n_batches=0)
        qlidar.compute_amax(bb, torch.device("cuda"))
        amax = {n[:-len(".act_quant")]: float(m.amax.detach().float().cpu().reshape(-1)[0]) for n, m in bb.named_modules()
                if n.endswith("act_quant")}
    eng = qlidar.BackboneEngine(bb, batch, coords.shape[0] + 700, use_graph=True, stage_cap_ratio=4.0, bev=False)
    for _ in range(2):                                                    # the second call replays the captured graph
        eng.forward_voxels(feats.cuda(), torch.from_numpy(coords).cuda())
    torch.cuda.synchronize()
    assert not eng.overflowed()
    counts = eng.counts()
    rec, out, taps = O.mirror_backbone_w8a8_pt(prog, P, feats, coords, O.sparse_shape_zyx(grid), batch, no_list=tuple(no_list), act_amax=amax)
    assert len(rec) == len(eng.layers)
    n_i8 = 0
    for L in eng.layers:
        r = rec[L.name]
        n_in, n_out = counts[L.stage_in], counts[L.stage_out]
        so = eng.stages[L.stage_out]
        assert n_out == r["out"].shape[0], L.name
        assert np.array_equal(so.coords[:n_out].cpu().numpy(), r["out_coords"]), L.name
        if L.kind == "i8":
            n_i8 += 1
            codes = L.q_buf[:n_in].cpu().numpy()
            bad = int((codes != r["codes"]).sum())
            assert bad == 0, f"{L.name}: {bad} of {codes.size} int8 codes differ"
            acc = torch.zeros((so.cap, L.cout), dtype=torch.int32, device="cuda")
            ops.spconv_mma(L.q_buf, eng.rulebooks[L.rb_key], so.cap, so.n_dev, L.cout, L.w, torch.ones(L.cout, device="cuda"),
                           torch.zeros(L.cout, device="cuda"), out=acc, kmask=eng.kmasks[L.rb_key], row_perm=eng.row_perms[L.rb_key])
            assert np.array_equal(acc[:n_out].cpu().numpy(), r["acc"]), f"{L.name}: INT32 accumulators differ"
        else:
            assert L.kind == "stem" and r["codes"] is None
        got = L.out[:n_out].cpu().numpy().view(np.uint16)
        bad = int((got != r["out"].view(np.uint16)).sum())
        assert bad == 0, f"{L.name}: {bad} of {got.size} fp16 outputs differ"
    assert n_i8 == len(eng.layers) - 1
    if static:
        assert sum(L.fused_q for L in eng.layers) == n_i8 - 1               # every quantised layer but the one fed by the stem
