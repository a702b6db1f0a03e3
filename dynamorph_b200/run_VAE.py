"""Drop-in for `python run_VAE.py -m process -c cfg.yml` (/root/reference/run_VAE.py:10-128).

One process per GPU, wells dealt round-robin, all GPUs busy at once (the reference starts one process per
well and joins it before starting the next, run_VAE.py:78-85).  Wells are independent: no collective."""
from __future__ import annotations

import argparse
import os
import re

import torch.multiprocessing as mp

from .configs.config_reader import YamlReader
from .pipeline.patch_VAE import process_VAE


def get_im_sites(input_dir):
    """`<WELL>-Site_<n>` identifiers present in a raw directory (reference: SingleCellPatch.extract_patches)."""
    sites = set()
    for f in os.listdir(input_dir):
        mt = re.match(r"^([A-Za-z0-9]+-Site_\d+)", f)
        if mt:
            sites.add(mt.group(1))
    wells = {f[:2] for f in os.listdir(input_dir) if f.endswith("_static_patches.pkl")}
    return sorted(sites) or sorted(w + "-Site_0" for w in wells)


def _worker(gpu, jobs, config_path):
    from .dist import bind_to_gpu_numa
    bind_to_gpu_numa(gpu)          # staging buffers of this GPU's encoder on its own NUMA node
    config = YamlReader().read_config(config_path)
    for raw_dir, supp_dir, well_sites in jobs:
        process_VAE(raw_dir, supp_dir, well_sites, config, gpu=gpu)


def main(method_, raw_dir_, supp_dir_, config_, config_path):
    if method_ != 'process':
        raise ValueError("only `-m process` is on the VQ-VAE hot path; assemble / trajectory_matching are CPU "
                         "pickle glue kept by the reference (SURVEY.md section 2)")
    if not raw_dir_:
        raise AttributeError("raw directory must be specified when method = process")
    if not config_.latent_encoding.weights:
        raise AttributeError("pytorch VQ-VAE weights path must be specified when method = process")
    gpus = list(getattr(config_.latent_encoding, "gpu_ids", [0]) or [0])
    sites = getattr(config_.latent_encoding, "fov", None) or get_im_sites(raw_dir_)
    wells = sorted(set(s[:2] for s in sites))
    per_gpu = {g: [] for g in gpus}
    for i, well in enumerate(wells):
        per_gpu[gpus[i % len(gpus)]].append((raw_dir_, supp_dir_, [s for s in sites if s[:2] == well]))
    mp.set_start_method('spawn', force=True)
    procs = [mp.Process(target=_worker, args=(g, jobs, config_path)) for g, jobs in per_gpu.items() if jobs]
    for p in procs:
        p.start()
    for p in procs:
        p.join()
        if p.exitcode != 0:
            raise RuntimeError(f"encoding worker exited with {p.exitcode}")


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
        main(arguments.method, raw_dir, supp_dir, config, arguments.config)
