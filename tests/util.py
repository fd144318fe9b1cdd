"""Synthetic suspensions shared by the tests and bench.py."""
import math

import numpy as np


def box_length(N, phi):
    return (4.0 * math.pi * N / (3.0 * phi)) ** (1.0 / 3.0)


def random_positions(N, L, seed=0):
    """i.i.d. uniform positions (overlaps allowed: exercises the RPY overlap branch)."""
    rng = np.random.default_rng(seed)
    pos = np.zeros((N, 4), dtype=np.float32)
    pos[:, :3] = (rng.random((N, 3)) - 0.5) * L
    return pos


def lattice_positions(N, L, seed=0, jitter=0.45):
    """Non-overlapping jittered FCC lattice filling the cubic box: first N sites of a seeded
    shuffle, each displaced by at most jitter * (surface gap) / sqrt(3) per axis."""
    rng = np.random.default_rng(seed)
    n = 1
    while 4 * n ** 3 < N:
        n += 1
    a = L / n
    base = np.array([[0, 0, 0], [0.5, 0.5, 0], [0.5, 0, 0.5], [0, 0.5, 0.5]])
    idx = np.stack(np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij"), -1).reshape(-1, 1, 3)
    sites = ((idx + base[None]) * a).reshape(-1, 3)
    sel = rng.permutation(len(sites))[:N]
    nn_dist = a / math.sqrt(2.0)
    gap = max(nn_dist - 2.0, 0.0)
    disp = (rng.random((N, 3)) * 2 - 1) * (jitter * gap / (2.0 * math.sqrt(3.0)))
    p = sites[sel] + disp - L / 2
    p = (p + L / 2) % L - L / 2
    pos = np.zeros((N, 4), dtype=np.float32)
    pos[:, :3] = p
    return pos


def random_forces(N, seed=1):
    rng = np.random.default_rng(seed)
    F = np.zeros((N, 4), dtype=np.float32)
    f = rng.standard_normal((N, 3))
    f -= f.mean(axis=0, keepdims=True)
    F[:, :3] = f
    return F


def rel_err(a, b):
    """(L2 relative error, max abs error / max |b|) over xyz."""
    a = np.asarray(a, dtype=np.float64)[:, :3]
    b = np.asarray(b, dtype=np.float64)[:, :3]
    return float(np.linalg.norm(a - b) / np.linalg.norm(b)), float(np.abs(a - b).max() / np.abs(b).max())
