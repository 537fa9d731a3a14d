"""Shared helpers for the parity tests (oracle <-> C-ABI layouts)."""
import numpy as np
import torch

TILE_M = 128


def nbr_to_tiles(nbr_kn: np.ndarray) -> np.ndarray:
    """oracle (K, N) -> device layout (tiles, K, 128), padded with -1."""
    K, N = nbr_kn.shape
    tiles = (N + TILE_M - 1) // TILE_M
    out = np.full((tiles, K, TILE_M), -1, dtype=np.int32)
    for t in range(tiles):
        n = min(TILE_M, N - t * TILE_M)
        out[t, :, :n] = nbr_kn[:, t * TILE_M:t * TILE_M + n]
    return out


def tiles_to_nbr(nbr_tiles: np.ndarray, n: int) -> np.ndarray:
    """device layout (tiles, K, 128) -> (K, n)."""
    tiles, K, _ = nbr_tiles.shape
    return np.ascontiguousarray(nbr_tiles.transpose(1, 0, 2).reshape(K, tiles * TILE_M)[:, :n])


def random_coords(rng, B, D, H, W, density):
    occ = rng.random((B, D, H, W)) < density
    c = np.argwhere(occ).astype(np.int32)
    return c[rng.permutation(len(c))]


def dense_nbr(ops, nbr, kmask, n: int) -> np.ndarray:
    """(K, n) oracle-form rulebook out of the kernels' COMPACT [tiles, K, 128] + per-tile mask form (include/qlidar.h)."""
    tiles = max(1, (int(n) + TILE_M - 1) // TILE_M)
    d = ops.expand_rulebook(nbr[:tiles], None if kmask is None else kmask[:tiles])
    return tiles_to_nbr(d.cpu().numpy(), int(n))
