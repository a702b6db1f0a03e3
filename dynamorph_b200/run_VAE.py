"""Drop-in for `python run_VAE.py -m process -c cfg.yml` (/root/reference/run_VAE.py:10-128).

One process per GPU, wells dealt round-robin, all GPUs busy at once (the reference starts one process per
well and joins it before starting the next, run_VAE.py:78-85).  Wells are independent: no collective."""
from __future__ import annotations

import argparse
import os

import torch.multiprocessing as mp

from .configs.config_reader import YamlReader
from .pipeline.patch_VAE import process_VAE


def get_im_sites(input_dir):
    """Site (field-of-view) names of a raw directory: the stems of its `.npy` image files, `_NN` probability maps
    excluded (reference: SingleCellPatch/extract_patches.py:337-350).  A directory that only holds assembled
    `<well>_static_patches.pkl` files yields one pseudo-site per well, which is all `process` needs."""
    stems = {os.path.splitext(f)[0] for f in os.listdir(input_dir) if f.endswith(".npy") and "_NN" not in f}
    if stems:
        return sorted(stems)
    return sorted(f[:-len("_static_patches.pkl")] + "-Site_0" for f in os.listdir(input_dir)
                  if f.endswith("_static_patches.pkl"))


def _worker(gpu, jobs, config_):
    import torch
    from .dist import bind_to_gpu_numa
    torch.cuda.set_device(gpu)     # everything this worker allocates or launches belongs to its GPU
    bind_to_gpu_numa(gpu)          # staging buffers of this GPU's encoder on its own NUMA node
    for raw_dir, supp_dir, well_sites in jobs:
        process_VAE(raw_dir, supp_dir, well_sites, config_, gpu=gpu)


def main(method_, raw_dir_, supp_dir_, config_, config_path=None):
    """run_VAE.py:28-93 for `method_ == 'process'`: wells of `raw_dir_` are dealt round-robin over
    `latent_encoding.gpu_ids`; each GPU's worker process encodes its wells one after another while the other GPUs
    work on theirs.  (`config_path` is accepted for older callers and ignored: the parsed config travels to the
    workers.)"""
    if method_ != 'process':
        raise ValueError("only `-m process` is on the VQ-VAE hot path; assemble / trajectory_matching are CPU "
                         "pickle glue kept by the reference (SURVEY.md section 2)")
    if not raw_dir_:
        raise AttributeError("raw directory must be specified when method = process")
    if not getattr(config_.latent_encoding, "weights", None):
        raise AttributeError("pytorch VQ-VAE weights path must be specified when method = process")
    from . import _lib
    _lib.load()                    # build (if needed) once here, not concurrently in every worker
    gpus = list(getattr(config_.latent_encoding, "gpu_ids", [0]) or [0])
    sites = getattr(config_.latent_encoding, "fov", None) or get_im_sites(raw_dir_)
    wells = sorted(set(s[:2] for s in sites))
    slots = [(g, []) for g in gpus]            # one worker per gpu_ids entry (an id listed twice gets two workers)
    for i, well in enumerate(wells):
        slots[i % len(slots)][1].append((raw_dir_, supp_dir_, [s for s in sites if s[:2] == well]))
    mp.set_start_method('spawn', force=True)
    procs = [mp.Process(target=_worker, args=(g, jobs, config_)) for g, jobs in slots if jobs]
    for p in procs:
        p.start()
    for p in procs:
        p.join()
    bad = [p.exitcode for p in procs if p.exitcode != 0]
    if bad:
        raise RuntimeError(f"encoding worker(s) exited with {bad}")


def parse_args():
    parser = argparse.ArgumentParser()
    parser.add_argument('-m', '--method', type=str, required=True,
                        choices=['assemble', 'process', 'trajectory_matching'], default='assemble',
                        help="Method: only 'process' is implemented here")
    parser.add_argument('-c', '--config', type=str, required=True, help='path to yaml configuration file')
    return parser.parse_args()


if __name__ == '__main__':
    arguments = parse_args()
    config = YamlReader().read_config(arguments.config)
    le = config.latent_encoding
    supp = getattr(le, "supp_dirs", None) or [None] * len(le.raw_dirs)
    for raw_dir, supp_dir in zip(le.raw_dirs, supp):
        main(arguments.method, raw_dir, supp_dir, config)
