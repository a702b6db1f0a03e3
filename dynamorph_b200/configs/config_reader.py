"""Minimal reader for the two config sections the VQ-VAE path consumes (`latent_encoding`, `training`),
same attribute access as /root/reference/configs/config_reader.py:140-206 (whose `yaml.load(f)` call no
longer works under PyYAML 6).  Unknown keys only warn, missing sections are empty namespaces."""
from __future__ import annotations

import logging
import types

import yaml

log = logging.getLogger(__name__)

LATENT_ENCODING = {'raw_dirs', 'supp_dirs', 'val_dirs', 'weights', 'save_output', 'gpu_ids', 'fov', 'channels',
                   'channel_mean', 'channel_std', 'num_classes', 'num_hiddens', 'num_residual_hiddens',
                   'num_embeddings', 'commitment_cost', 'network', 'patch_type', 'w_a', 'w_t', 'margin'}
TRAINING = {'raw_dirs', 'supp_dirs', 'weights_dirs', 'network', 'num_inputs', 'num_hiddens', 'num_residual_hiddens',
            'num_residual_layers', 'num_embeddings', 'weight_matching', 'margin', 'w_a', 'w_t', 'w_n', 'channel_mean', 'channel_std',
            'commitment_cost', 'n_epochs', 'learn_rate', 'batch_size', 'val_split_ratio', 'shuffle_data',
            'transform', 'patience', 'n_pos_samples', 'num_workers', 'gpu_id', 'start_model_path', 'retrain',
            'start_epoch', 'earlystop_metric', 'model_name', 'use_mask', 'channels', 'temperature', 'augmentations'}


class YamlReader:
    def __init__(self):
        self.config = None
        self.latent_encoding = types.SimpleNamespace()
        self.training = types.SimpleNamespace()

    def read_config(self, yml_config):
        with open(yml_config, 'r') as f:
            self.config = yaml.safe_load(f) or {}
        for section, allowed in (('latent_encoding', LATENT_ENCODING), ('training', TRAINING)):
            ns = getattr(self, section)
            for key, value in (self.config.get(section) or {}).items():
                if key not in allowed:
                    log.warning(f"yaml {section} config field {key} is not recognized")
                setattr(ns, key, value)
        return self
