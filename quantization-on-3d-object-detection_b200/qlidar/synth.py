"""Frozen synthetic LiDAR generators ("G1 lidar-like", "G2 surface sheet") and the dataset
constants of the reference's yaml files.  Pure numpy; used by bench.py and the examples.  tests/test_synth.py checks
that the oracle's copy produces identical arrays."""
from __future__ import annotations

from typing import List, Optional

import numpy as np

CONFIGS = {
    # name: (point_cloud_range xyzxyz, voxel_size xyz, num point features, max pts/voxel, max voxels @test)
    "kitti": dict(pc_range=[0.0, -40.0, -3.0, 70.4, 40.0, 1.0], voxel_size=[0.05, 0.05, 0.1], nfeat=4,
                  max_pts=5, max_voxels=40000),        # kitti_dataset.yaml:4,50,65-70
    "waymo": dict(pc_range=[-75.2, -75.2, -2.0, 75.2, 75.2, 4.0], voxel_size=[0.1, 0.1, 0.15], nfeat=5,
                  max_pts=5, max_voxels=150000),       # waymo_dataset.yaml:5,63,79-84
    "nuscenes": dict(pc_range=[-54.0, -54.0, -5.0, 54.0, 54.0, 3.0], voxel_size=[0.075, 0.075, 0.2], nfeat=5,
                     max_pts=10, max_voxels=160000),   # cbgs_voxel0075_res3d_centerpoint.yaml:6,55-60
}


def grid_size_xyz(pc_range, voxel_size) -> np.ndarray:
    """pcdet/datasets/processor/data_processor.py:135-137: round((max-min)/voxel) as int64, xyz order."""
    r = np.asarray(pc_range, dtype=np.float64)
    g = (r[3:6] - r[0:3]) / np.asarray(voxel_size, dtype=np.float64)
    return np.round(g).astype(np.int64)


def sparse_shape_zyx(grid_xyz) -> List[int]:
    g = [int(v) for v in grid_xyz]
    return [g[2] + 1, g[1], g[0]]


def synth_lidar_frame(cfg: str, seed: int, n_az: Optional[int] = None, n_beams: Optional[int] = None) -> np.ndarray:
    """G1: sensor at origin above a ground plane, `n_beams` elevation rings x `n_az` azimuth steps; each ray hits
    the ground or one of a set of random vertical cylinders; range noise N(0,0.02). Returns (P, F) float32."""
    rng = np.random.default_rng(seed)
    if cfg == "kitti":
        elev = np.deg2rad(np.linspace(-24.8, 2.0, n_beams or 64)); n_az = n_az or 1400
        sensor_h, n_cyl, ext, max_r, nf = 1.73, 60, 60.0, 80.0, 4
    else:
        elev = np.deg2rad(np.linspace(-17.6, 2.4, n_beams or 192)); n_az = n_az or 2650
        sensor_h, n_cyl, ext, max_r, nf = 2.0, 120, 75.0, 75.0, 5
    az = np.linspace(-np.pi, np.pi, n_az, endpoint=False)
    cyl_c = rng.uniform(-ext, ext, size=(n_cyl, 2))
    cyl_r = rng.uniform(0.5, 2.5, size=n_cyl)
    cyl_h = rng.uniform(1.5, 3.5, size=n_cyl)
    A, E = np.meshgrid(az, elev, indexing="ij")
    dx, dy, dz = np.cos(E) * np.cos(A), np.cos(E) * np.sin(A), np.sin(E)
    with np.errstate(divide="ignore", invalid="ignore"):
        t_ground = np.where(dz < -1e-6, -sensor_h / dz, np.inf)
    t_best = np.minimum(t_ground, max_r * 1.5)
    dxy2 = dx * dx + dy * dy
    for c, r, h in zip(cyl_c, cyl_r, cyl_h):
        # ray-circle intersection in the xy plane
        b = dx * c[0] + dy * c[1]
        cc = c[0] * c[0] + c[1] * c[1] - r * r
        disc = b * b - dxy2 * cc
        with np.errstate(invalid="ignore"):
            t = (b - np.sqrt(np.where(disc >= 0, disc, np.nan))) / dxy2
        z = t * dz                                    # relative to the sensor
        ok = (disc >= 0) & (t > 0.5) & (z >= -sensor_h) & (z <= -sensor_h + h)
        t_best = np.where(ok & (t < t_best), t, t_best)
    t_best = t_best + rng.normal(0.0, 0.02, size=t_best.shape)
    valid = np.isfinite(t_best) & (t_best < max_r) & (t_best > 0.5)
    x, y, z = (t_best * dx)[valid], (t_best * dy)[valid], (t_best * dz)[valid]
    if cfg != "kitti":
        z = z + sensor_h                              # vehicle frame: ground at z = 0 (KITTI stays in the sensor frame)
    if cfg == "kitti":
        keep = (x > 0) & (np.abs(np.arctan2(y, x)) < np.pi / 4)
        x, y, z = x[keep], y[keep], z[keep]
    feats = [x, y, z] + [rng.uniform(0, 1, size=x.shape) for _ in range(nf - 3)]
    pts = np.stack(feats, axis=1).astype(np.float32)
    if cfg != "kitti":                                # Waymo/nuScenes shuffle points at test time (waymo_dataset.yaml:72-76)
        pts = pts[rng.permutation(pts.shape[0])]
    return pts


def synth_batch(cfg: str, batch: int, first_seed: int = 1000, **kw) -> np.ndarray:
    """Collated `points (sum P, 1+F)` with the batch index in column 0 (dataset.py collate of 'points')."""
    fr = [synth_lidar_frame(cfg, first_seed + i, **kw) for i in range(batch)]
    return np.concatenate([np.concatenate([np.full((f.shape[0], 1), i, np.float32), f], axis=1) for i, f in enumerate(fr)])


def synth_surface_sheet(S: int, seed: int = 2000, depth: int = 40) -> np.ndarray:
    """G2: S x S (x,y) patch with z0(x,y) a clipped 2-D random walk in [0,depth): N = S^2 voxels. Returns (N,4) [b,z,y,x]."""
    rng = np.random.default_rng(seed)
    steps_y = rng.integers(-1, 2, size=(S, 1)).cumsum(axis=0)
    steps_x = rng.integers(-1, 2, size=(S, S)).cumsum(axis=1)
    z = np.clip(depth // 2 + steps_y + steps_x, 0, depth - 1)
    yy, xx = np.meshgrid(np.arange(S), np.arange(S), indexing="ij")
    return np.stack([np.zeros(S * S, np.int64), z.ravel(), yy.ravel(), xx.ravel()], axis=1).astype(np.int32)
