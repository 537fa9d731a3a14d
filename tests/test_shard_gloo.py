"""Multi-process host logic of the frame-sharded path on CPU: world_size 2, gloo, 127.0.0.1.
Mirrors pcdet/datasets/__init__.py:31-52 (sampler) and pcdet/utils/common_utils.py:229-250 (merge)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_frames_for_rank_matches_reference_sampler():
    from qlidar import shard
    for n, w in [(64, 8), (10, 4), (7, 2), (3, 4), (1, 1)]:
        per = (n + w - 1) // w
        seen = []
        for r in range(w):
            idx = shard.frames_for_rank(n, r, w)
            assert len(idx) == per
            # torch.utils.data.DistributedSampler arithmetic restated: pad by wrap-around, stride by world
            ref = (list(range(n)) + list(range(n))[:per * w - n])[r:per * w:w]
            assert idx == ref
            seen.append(idx)
        order = shard.merge_order(n, w)
        assert [seen[r][j] for r, j in order] == list(range(n))          # the merge restores dataset order
    with pytest.raises(ValueError):
        shard.frames_for_rank(4, 2, 2)


def _worker(rank, world, port, n_frames, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from qlidar import shard
    mine = shard.frames_for_rank(n_frames, rank, world)
    # per-frame "detections": a padded [max_det, 9] block whose content is a function of the dataset frame index
    local = torch.stack([torch.full((5, 9), float(f)) + torch.arange(9.0) for f in mine])
    merged = shard.gather_frame_results(local, n_frames)
    ok = merged.shape == (n_frames, 5, 9) and all(torch.equal(merged[i], torch.full((5, 9), float(i)) + torch.arange(9.0)) for i in range(n_frames))
    # timing reduction used by bench.py: max over ranks
    t = torch.tensor([1.0 + rank])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    rate = shard.aggregate_rate([len(mine)] * world, float(t.item()))
    q.put((rank, bool(ok), rate))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_frames", [8, 7])
def test_gather_frame_results_world2_gloo(n_frames):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    per = (n_frames + world - 1) // world
    for rank, ok, rate in res:
        assert ok, f"rank {rank}: merged results out of order"
        assert abs(rate - (per * world) / 2.0) < 1e-9                   # all ranks' frames / slowest rank's time (2.0 s)


def test_single_process_passthrough():
    from qlidar import shard
    x = torch.arange(12.0).reshape(4, 3)
    assert torch.equal(shard.gather_frame_results(x, 3), x[:3])


def test_bench_head_maps_are_a_function_of_the_dataset_frame_index():
    """bench.py's detections gate compares what the ranks gathered with a one-process recomputation: the synthetic head maps of a
    frame must depend on its dataset index alone (not on the rank, the batch it is stacked into or the call order)."""
    import importlib.util
    import numpy as np
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_for_test", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    a, b, c = bench.frame_head_maps(5), bench.frame_head_maps(6), bench.frame_head_maps(5)
    assert set(a) == {"hm", "center", "center_z", "dim", "rot"} and a["hm"].shape == (3, 188, 188)
    assert all(np.array_equal(a[k], c[k]) for k in a) and not np.array_equal(a["hm"], b["hm"])
    assert (1.0 / (1.0 + np.exp(-a["hm"])) > 0.1).sum() > 300                 # enough candidates above SCORE_THRESH for the NMS to work on
