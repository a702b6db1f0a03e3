"""Sharded latent store for multi-million-patch runs.

The reference writes ONE pickle per array, `<well>_latent_space.pkl` / `_latent_space_after.pkl`, each an
`np.float32 (N, D*h*w)` NCHW-flattened matrix (/root/reference/pipeline/patch_VAE.py:454-462).  At 4 M patches that
is 2 x 65.5 GB: more than the host RAM of the box and nothing one process can pickle.  Here every rank appends
fixed-size `.npy` shards of its own patch range (rows keep the reference's layout bit for bit, so a pre-fitted
`pca_model.pkl` still applies) and writes a small JSON manifest; `open_latents` gives back a read-only view that
behaves like the (N, L) array (row slicing, `__len__`, chunk iteration, `to_array()` when it fits), and
`write_reference_pickle` re-creates the reference's single-pickle file for small runs.

    w = ShardedLatentWriter(out_dir, "B2", "latent_space", width=4096, rank=r, world=R, rows_per_shard=65536)
    for chunk in ...: w.append(z_before_chunk)          # (n, 4096) float32, host
    w.close()                                            # rank-local manifest
    merge_manifests(out_dir, "B2", "latent_space")      # once, after all ranks closed (rank 0)
    z = open_latents(out_dir, "B2", "latent_space")     # z[a:b] -> np.ndarray
"""
from __future__ import annotations

import json
import os
import pickle
from typing import Iterator, List, Optional, Tuple

import numpy as np

FORMAT = "dynamorph_b200.latent_shards.v1"


def _shard_name(well: str, kind: str, rank: int, index: int) -> str:
    return f"{well}_{kind}.r{rank:03d}.s{index:05d}.npy"


def _manifest_name(well: str, kind: str, rank: Optional[int] = None) -> str:
    return f"{well}_{kind}.manifest.json" if rank is None else f"{well}_{kind}.r{rank:03d}.manifest.json"


class ShardedLatentWriter:
    """Append-only writer of one rank's rows.  `first_row` is the global index of this rank's first patch
    (dist.shard_range(N, rank, world)[0]); rows are buffered until a shard is full, so memory stays at one shard."""

    def __init__(self, out_dir: str, well: str, kind: str, width: int, rank: int = 0, world: int = 1,
                 first_row: int = 0, rows_per_shard: int = 65536, dtype=np.float32, run_id: Optional[str] = None):
        if rows_per_shard <= 0 or width <= 0:
            raise ValueError("rows_per_shard and width must be positive")
        if not (0 <= rank < world):
            raise ValueError(f"rank {rank} outside [0, {world})")
        self.out_dir, self.well, self.kind = out_dir, well, kind
        self.width, self.rank, self.world = int(width), int(rank), int(world)
        self.first_row, self.rows_per_shard = int(first_row), int(rows_per_shard)
        self.dtype = np.dtype(dtype)
        self.run_id = run_id
        os.makedirs(out_dir, exist_ok=True)
        # a re-run into the same directory: this rank's manifest and shards of the earlier run go first (manifest
        # before shards, so that a reader never sees a manifest whose shards are gone), and so does the merged manifest
        stem = f"{well}_{kind}.r{self.rank:03d}."
        old = [f for f in os.listdir(out_dir) if f.startswith(stem)]
        for f in sorted(old, key=lambda f: not f.endswith(".manifest.json")) + [_manifest_name(well, kind)]:
            try:
                os.remove(os.path.join(out_dir, f))
            except FileNotFoundError:
                pass
        self._buf = np.empty((self.rows_per_shard, self.width), dtype=self.dtype)
        self._fill = 0
        self._rows = 0
        self._shards: List[dict] = []
        self._closed = False

    def append(self, rows) -> None:
        if self._closed:
            raise RuntimeError("writer is closed")
        a = np.asarray(rows.numpy() if hasattr(rows, "numpy") else rows)
        if a.ndim != 2 or a.shape[1] != self.width:
            raise ValueError(f"expected (n, {self.width}) rows, got {a.shape}")
        if a.dtype != self.dtype:
            raise ValueError(f"expected {self.dtype} rows, got {a.dtype}")
        pos = 0
        while pos < a.shape[0]:
            n = min(a.shape[0] - pos, self.rows_per_shard - self._fill)
            self._buf[self._fill:self._fill + n] = a[pos:pos + n]
            self._fill += n
            pos += n
            if self._fill == self.rows_per_shard:
                self._flush()

    def _flush(self) -> None:
        if self._fill == 0:
            return
        name = _shard_name(self.well, self.kind, self.rank, len(self._shards))
        tmp = os.path.join(self.out_dir, name + ".tmp")
        with open(tmp, "wb") as f:
            np.save(f, self._buf[:self._fill])
        os.replace(tmp, os.path.join(self.out_dir, name))        # a reader never sees a half-written shard
        self._shards.append({"file": name, "first_row": self.first_row + self._rows, "rows": int(self._fill)})
        self._rows += self._fill
        self._fill = 0

    def close(self) -> str:
        if not self._closed:
            self._flush()
            man = {"format": FORMAT, "well": self.well, "kind": self.kind, "width": self.width,
                   "dtype": self.dtype.str, "rank": self.rank, "world": self.world, "run_id": self.run_id,
                   "first_row": self.first_row,
                   "rows": self._rows, "shards": self._shards}
            path = os.path.join(self.out_dir, _manifest_name(self.well, self.kind, self.rank))
            with open(path + ".tmp", "w") as f:
                json.dump(man, f)
            os.replace(path + ".tmp", path)
            self._closed = True
        return os.path.join(self.out_dir, _manifest_name(self.well, self.kind, self.rank))

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def merge_manifests(out_dir: str, well: str, kind: str) -> str:
    """Combine the per-rank manifests into `<well>_<kind>.manifest.json`; checks that the ranks are all there and
    that their row ranges tile [0, N) without gap or overlap.  Rank manifests left behind by an EARLIER run with
    another world size or run id (ranks that no longer exist) are ignored: the newest manifest names the run."""
    parts = []
    for fn in sorted(os.listdir(out_dir)):
        if fn.startswith(f"{well}_{kind}.r") and fn.endswith(".manifest.json"):
            path = os.path.join(out_dir, fn)
            with open(path) as f:
                parts.append((os.path.getmtime(path), json.load(f)))
    if not parts:
        raise FileNotFoundError(f"no rank manifests for {well}_{kind} in {out_dir}")
    newest = max(parts, key=lambda p: p[0])[1]
    parts = [p for _, p in parts if p["world"] == newest["world"] and p.get("run_id") == newest.get("run_id")]
    world = parts[0]["world"]
    ranks = sorted(p["rank"] for p in parts)
    if ranks != list(range(world)):
        raise ValueError(f"rank manifests {ranks} do not cover world size {world}")
    if len({(p["width"], p["dtype"], p["format"]) for p in parts}) != 1:
        raise ValueError("rank manifests disagree on width / dtype / format")
    shards = sorted((s for p in parts for s in p["shards"]), key=lambda s: s["first_row"])
    row = 0
    for s in shards:
        if s["first_row"] != row:
            raise ValueError(f"shard {s['file']} starts at row {s['first_row']}, expected {row}")
        row += s["rows"]
    man = {"format": FORMAT, "well": well, "kind": kind, "width": parts[0]["width"], "dtype": parts[0]["dtype"],
           "world": world, "rows": row, "shards": shards}
    path = os.path.join(out_dir, _manifest_name(well, kind))
    with open(path + ".tmp", "w") as f:
        json.dump(man, f)
    os.replace(path + ".tmp", path)
    return path


class ShardedLatents:
    """Read-only (N, L) view over the shards: len(), shape, dtype, z[i], z[a:b] (contiguous slices), iteration in
    shard-sized chunks.  Shards are memory-mapped, so a slice only touches the files it covers."""

    def __init__(self, out_dir: str, manifest: dict):
        if manifest.get("format") != FORMAT:
            raise ValueError(f"not a {FORMAT} manifest")
        self.out_dir = out_dir
        self.manifest = manifest
        self.shape: Tuple[int, int] = (int(manifest["rows"]), int(manifest["width"]))
        self.dtype = np.dtype(manifest["dtype"])
        self._starts = np.asarray([s["first_row"] for s in manifest["shards"]], dtype=np.int64)
        self._maps: dict = {}

    def __len__(self) -> int:
        return self.shape[0]

    def _shard(self, i: int) -> np.ndarray:
        m = self._maps.get(i)
        if m is None:
            s = self.manifest["shards"][i]
            m = np.load(os.path.join(self.out_dir, s["file"]), mmap_mode="r")
            if m.shape != (s["rows"], self.shape[1]) or m.dtype != self.dtype:
                raise ValueError(f"shard {s['file']} is {m.shape} {m.dtype}, manifest says ({s['rows']}, {self.shape[1]})")
            self._maps[i] = m
        return m

    def __getitem__(self, key) -> np.ndarray:
        if isinstance(key, (int, np.integer)):
            k = int(key) + (self.shape[0] if key < 0 else 0)
            if not 0 <= k < self.shape[0]:
                raise IndexError(key)
            return self[k:k + 1][0]
        if not isinstance(key, slice):
            raise TypeError("ShardedLatents supports integer and contiguous slice indexing")
        a, b, step = key.indices(self.shape[0])
        if step != 1:
            raise TypeError("ShardedLatents supports contiguous slices only")
        out = np.empty((max(0, b - a), self.shape[1]), dtype=self.dtype)
        if b <= a:
            return out
        i = int(np.searchsorted(self._starts, a, side="right") - 1)
        pos = a
        while pos < b:
            s = self.manifest["shards"][i]
            lo = pos - s["first_row"]
            n = min(b - pos, s["rows"] - lo)
            out[pos - a:pos - a + n] = self._shard(i)[lo:lo + n]
            pos += n
            i += 1
        return out

    def chunks(self) -> Iterator[Tuple[int, np.ndarray]]:
        for i, s in enumerate(self.manifest["shards"]):
            yield s["first_row"], np.asarray(self._shard(i))

    def to_array(self) -> np.ndarray:
        return self[0:self.shape[0]]


def open_latents(out_dir: str, well: str, kind: str) -> ShardedLatents:
    with open(os.path.join(out_dir, _manifest_name(well, kind))) as f:
        return ShardedLatents(out_dir, json.load(f))


def write_reference_pickle(out_dir: str, well: str, kind: str) -> str:
    """`<well>_<kind>.pkl` exactly as the reference writes it (np.float32 (N, L), pickle protocol 4,
    patch_VAE.py:456-461) -- for runs small enough to hold in memory."""
    z = open_latents(out_dir, well, kind).to_array()
    path = os.path.join(out_dir, f"{well}_{kind}.pkl")
    with open(path, "wb") as f:
        pickle.dump(z, f, protocol=4)
    return path
