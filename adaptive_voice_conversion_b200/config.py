"""Hyper-parameter dictionaries in the reference's config.yaml schema (config.yaml:1-52).

``default_config(c_in)`` is the shipped configuration with the mel dimension overridable:
BASELINE.json benchmarks 80 mels, the reference's own config.yaml uses 512.
"""
from __future__ import annotations

import copy

import yaml

_BASE = yaml.safe_load("""
SpeakerEncoder: {c_in: 512, c_h: 128, c_out: 128, kernel_size: 5, bank_size: 8, bank_scale: 1, c_bank: 128,
                 n_conv_blocks: 6, n_dense_blocks: 6, subsample: [1, 2, 1, 2, 1, 2], act: relu, dropout_rate: 0}
ContentEncoder: {c_in: 512, c_h: 128, c_out: 128, kernel_size: 5, bank_size: 8, bank_scale: 1, c_bank: 128,
                 n_conv_blocks: 6, subsample: [1, 2, 1, 2, 1, 2], act: relu, dropout_rate: 0}
Decoder: {c_in: 128, c_cond: 128, c_h: 128, c_out: 512, kernel_size: 5, n_conv_blocks: 6,
          upsample: [2, 1, 2, 1, 2, 1], act: relu, sn: false, dropout_rate: 0}
data_loader: {segment_size: 128, frame_size: 1, batch_size: 128, shuffle: true}
optimizer: {lr: 0.0005, beta1: 0.9, beta2: 0.999, amsgrad: true, weight_decay: 0.0001, grad_norm: 5}
lambda: {lambda_rec: 10, lambda_kl: 1}
annealing_iters: 20000
""")


def default_config(c_in: int = 512) -> dict:
    cfg = copy.deepcopy(_BASE)
    cfg["SpeakerEncoder"]["c_in"] = c_in
    cfg["ContentEncoder"]["c_in"] = c_in
    cfg["Decoder"]["c_out"] = c_in
    return cfg


def load_config(path: str) -> dict:
    """yaml.safe_load (the reference's bare yaml.load(f), main.py:28, fails on PyYAML >= 6)."""
    with open(path) as f:
        return yaml.safe_load(f)
