"""Shared helpers for the parity tests."""
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    path = os.path.join(GOLDEN, name + ".npz")
    z = np.load(path, allow_pickle=False)
    return {k: z[k] for k in z.files}


def sampled_rel_l2(t: torch.Tensor, g: dict, key: str) -> float:
    """rel-L2 between tensor ``t`` and the golden SAMPLE stored under ``key`` (e.g. 'out.res2')."""
    idx = torch.from_numpy(g[key + ".idx"])
    ref = torch.from_numpy(g[key + ".val"]).double()
    assert tuple(t.shape) == tuple(int(s) for s in g[key + ".shape"]), (t.shape, g[key + ".shape"])
    got = t.detach().reshape(-1).cpu().double()[idx]
    return float((got - ref).norm() / ref.norm())


def norm_ratio(t: torch.Tensor, g: dict, key: str) -> float:
    return float(t.detach().double().norm().item() / float(g[key + ".norm"]))
