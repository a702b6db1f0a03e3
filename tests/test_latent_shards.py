"""CPU: the sharded latent store (dynamorph_b200/latent_shards.py) round-trips the reference's
(N, D*h*w) float32 layout (pipeline/patch_VAE.py:454-462) across ranks, ragged appends and shard boundaries."""
import os
import pickle

import numpy as np
import pytest

from dynamorph_b200.dist import shard_range
from dynamorph_b200.latent_shards import (ShardedLatentWriter, merge_manifests, open_latents, write_reference_pickle)


def _write(tmp, z, world, rows_per_shard, pieces):
    N = z.shape[0]
    for r in range(world):
        a, b = shard_range(N, r, world)
        with ShardedLatentWriter(str(tmp), "B2", "latent_space", z.shape[1], rank=r, world=world, first_row=a,
                                 rows_per_shard=rows_per_shard) as w:
            pos = a
            for n in pieces:
                n = min(n, b - pos)
                if n > 0:
                    w.append(z[pos:pos + n])
                    pos += n
            if pos < b:
                w.append(z[pos:b])
    return merge_manifests(str(tmp), "B2", "latent_space")


@pytest.mark.parametrize("world,rows_per_shard", [(1, 64), (3, 7), (4, 1000)])
def test_round_trip_matches_reference_layout(tmp_path, world, rows_per_shard):
    rng = np.random.default_rng(0)
    z = rng.standard_normal((101, 48)).astype(np.float32)
    _write(tmp_path, z, world, rows_per_shard, pieces=[1, 5, 13, 2])
    v = open_latents(str(tmp_path), "B2", "latent_space")
    assert v.shape == z.shape and len(v) == 101 and v.dtype == np.float32
    assert np.array_equal(v.to_array(), z)
    assert np.array_equal(v[17:63], z[17:63]) and np.array_equal(v[-1], z[-1]) and v[5:5].shape == (0, 48)
    rows = 0
    for first, chunk in v.chunks():
        assert np.array_equal(chunk, z[first:first + chunk.shape[0]])
        rows += chunk.shape[0]
    assert rows == 101
    p = write_reference_pickle(str(tmp_path), "B2", "latent_space")
    assert os.path.basename(p) == "B2_latent_space.pkl"           # the reference's file name
    with open(p, "rb") as f:
        back = pickle.load(f)
    assert back.dtype == np.float32 and np.array_equal(back, z)


def test_missing_rank_and_bad_rows_are_errors(tmp_path):
    z = np.zeros((10, 4), np.float32)
    with ShardedLatentWriter(str(tmp_path), "C3", "latent_space_after", 4, rank=0, world=2, first_row=0) as w:
        w.append(z)
        with pytest.raises(ValueError):
            w.append(np.zeros((3, 5), np.float32))
        with pytest.raises(ValueError):
            w.append(np.zeros((3, 4), np.float64))
    with pytest.raises(ValueError):
        merge_manifests(str(tmp_path), "C3", "latent_space_after")       # rank 1 never wrote
    with ShardedLatentWriter(str(tmp_path), "C3", "latent_space_after", 4, rank=1, world=2, first_row=11) as w:
        w.append(z)
    with pytest.raises(ValueError):
        merge_manifests(str(tmp_path), "C3", "latent_space_after")       # gap between the ranks' ranges
    with pytest.raises(RuntimeError):
        w.append(z)


def test_empty_rank_is_fine(tmp_path):
    z = np.arange(8, dtype=np.float32).reshape(2, 4)
    for r in range(4):                                                   # more ranks than rows
        a, b = shard_range(2, r, 4)
        with ShardedLatentWriter(str(tmp_path), "D4", "latent_space", 4, rank=r, world=4, first_row=a) as w:
            if b > a:
                w.append(z[a:b])
    merge_manifests(str(tmp_path), "D4", "latent_space")
    assert np.array_equal(open_latents(str(tmp_path), "D4", "latent_space").to_array(), z)


def test_rerun_into_same_directory_ignores_earlier_world(tmp_path):
    """process_VAE re-runs into the same output directory: rank files of an earlier run with MORE ranks must not be
    merged in, and this run's ranks replace their own earlier shards."""
    rng = np.random.default_rng(1)
    old = rng.standard_normal((40, 8)).astype(np.float32)
    new = rng.standard_normal((23, 8)).astype(np.float32)
    _write(tmp_path, old, 4, 5, pieces=[3, 9])
    _write(tmp_path, new, 2, 50, pieces=[7])
    v = open_latents(str(tmp_path), "B2", "latent_space")
    assert np.array_equal(v.to_array(), new)
    # ranks 0 and 1 dropped every shard of the earlier run (s00001.. would otherwise linger beside the manifest)
    left = sorted(f for f in os.listdir(tmp_path) if ".r000." in f or ".r001." in f)
    assert left == ["B2_latent_space.r000.manifest.json", "B2_latent_space.r000.s00000.npy",
                    "B2_latent_space.r001.manifest.json", "B2_latent_space.r001.s00000.npy"]
