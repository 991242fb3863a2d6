"""Minimal stand-ins for the reference's data_utils.py so Solver keeps its constructor
contract.  The data pipeline is out of scope: ``PickleDataset``
/ ``get_data_loader`` read the same pickle + index-json formats (data_utils.py:43-57,
10-28) with the stock DataLoader, and ``SyntheticSegments`` provides the N(0,1) segments
BASELINE.json benchmarks on.
"""
from __future__ import annotations

import json
import pickle

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset


class PickleDataset(Dataset):
    """(utt_id, t) index over a dict of [T, n_mels] arrays -> [segment_size, n_mels] crops."""

    def __init__(self, pickle_path, sample_index_path, segment_size):
        with open(pickle_path, "rb") as f:
            self.data = pickle.load(f)
        with open(sample_index_path) as f:
            self.indexes = json.load(f)
        self.segment_size = segment_size

    def __len__(self):
        return len(self.indexes)

    def __getitem__(self, i):
        utt, t = self.indexes[i]
        return self.data[utt][t:t + self.segment_size]


class CollateFn:
    """[B, T, n_mels] crops -> [B, n_mels*frame_size, T/frame_size] (data_utils.py:10-22)."""

    def __init__(self, frame_size):
        self.frame_size = frame_size

    def __call__(self, items):
        t = torch.from_numpy(np.asarray(items, dtype=np.float32))
        b, n, m = t.shape
        return t.reshape(b, n // self.frame_size, self.frame_size * m).transpose(1, 2).contiguous()


def get_data_loader(dataset, batch_size, frame_size, shuffle=True, num_workers=4, drop_last=False):
    return DataLoader(dataset, batch_size=batch_size, shuffle=shuffle, num_workers=num_workers,
                      collate_fn=CollateFn(frame_size), pin_memory=True, drop_last=drop_last)


class SyntheticSegments:
    """Endless iterator of pinned N(0,1) batches [B, n_mels, T] (training data is per-mel
    z-normalised, so N(0,1) is representative)."""

    def __init__(self, batch_size, n_mels, segment_size, seed=1, n_distinct=4):
        g = torch.Generator().manual_seed(seed)
        self.batches = [torch.randn((batch_size, n_mels, segment_size), generator=g) for _ in range(n_distinct)]
        if torch.cuda.is_available():
            self.batches = [b.pin_memory() for b in self.batches]
        self.i = 0

    def __iter__(self):
        return self

    def __next__(self):
        b = self.batches[self.i % len(self.batches)]
        self.i += 1
        return b
