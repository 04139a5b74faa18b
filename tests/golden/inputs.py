"""Seeded synthetic inputs shared by the golden generator (make_golden.py) and the tests.

Inputs are regenerated from their seeds on both boxes (numpy Generator streams are stable), so the
committed .npz files only carry the reference's OUTPUTS.
"""
from __future__ import annotations

import numpy as np

ARCFACE_TEMPLATE = np.array([[38.2946, 51.6963], [73.5318, 51.5014], [56.0252, 71.7366],
                             [41.5493, 92.3655], [70.7299, 92.2041]], dtype=np.float32)


def head_tensors(seed: int, in_h: int = 640, in_w: int = 640, tie_fraction: float = 0.0, score_mu: float = -4.0):
    """Nine SCRFD outputs in reference order/shape: scores sigmoid(N(mu,1.5)), bbox U(0.5,6), kps N(0,2)."""
    rng = np.random.default_rng(seed)
    scores, bboxes, kpss = [], [], []
    for s in (8, 16, 32):
        n = (in_h // s) * (in_w // s) * 2
        z = rng.normal(score_mu, 1.5, (n, 1))
        sc = (1.0 / (1.0 + np.exp(-z))).astype(np.float32)
        if tie_fraction > 0:                       # duplicate some scores to exercise the tie order
            k = int(n * tie_fraction)
            src = rng.integers(0, n, k)
            dst = rng.integers(0, n, k)
            sc[dst] = sc[src]
        scores.append(sc)
        bboxes.append(rng.uniform(0.5, 6.0, (n, 4)).astype(np.float32))
        kpss.append(rng.normal(0.0, 2.0, (n, 10)).astype(np.float32))
    return scores + bboxes + kpss


def frame(seed: int, h: int, w: int) -> np.ndarray:
    return np.random.default_rng(seed).integers(0, 256, (h, w, 3), dtype=np.uint8)


def smooth_frame(seed: int, h: int, w: int) -> np.ndarray:
    """Low-frequency image (so interpolation errors are visible, unlike white noise)."""
    import cv2
    rng = np.random.default_rng(seed)
    small = rng.integers(0, 256, (h // 16 + 2, w // 16 + 2, 3), dtype=np.uint8)
    return cv2.resize(small, (w, h), interpolation=cv2.INTER_CUBIC)


def landmarks(seed: int, h: int, w: int, count: int) -> np.ndarray:
    """`count` plausible five-point sets (rotated / scaled template + jitter), float32 [count,5,2]."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(count):
        c = rng.uniform([0.1 * w, 0.1 * h], [0.9 * w, 0.9 * h])
        s = rng.uniform(0.3, 3.0)
        th = rng.uniform(-0.6, 0.6)
        R = s * np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
        out.append(((ARCFACE_TEMPLATE - 56.0) @ R.T + c + rng.normal(0, 1.5, (5, 2))).astype(np.float32))
    return np.stack(out)


def embeddings(seed: int, n: int, dim: int = 512) -> np.ndarray:
    return np.random.default_rng(seed).standard_normal((n, dim)).astype(np.float32)


def planted_queries(gallery: np.ndarray, seed: int, q: int, noise: float = 1.0):
    """Queries = gallery rows at random ids + noise (cos to the planted row ~0.7); returns (queries, ids)."""
    rng = np.random.default_rng(seed)
    ids = rng.integers(0, len(gallery), q)
    g = gallery[ids] / np.linalg.norm(gallery[ids], axis=1, keepdims=True)
    n = rng.standard_normal(g.shape).astype(np.float32)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    return (g + noise * n).astype(np.float32) * rng.uniform(5, 30, (q, 1)).astype(np.float32), ids


def clustered(seed: int, centres: int, members: int, dim: int = 512, noise: float = 0.35) -> np.ndarray:
    """centres x members noisy copies, shuffled; pairwise cos within a cluster ~ 1/(1+noise^2)."""
    rng = np.random.default_rng(seed)
    c = rng.standard_normal((centres, dim)).astype(np.float32)
    c /= np.linalg.norm(c, axis=1, keepdims=True)
    x = np.repeat(c, members, axis=0)
    n = rng.standard_normal(x.shape).astype(np.float32)
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    x = x + noise * n
    return x[rng.permutation(len(x))].astype(np.float32)


def chains(seed: int, n_chains: int, length: int, step: float, dim: int = 512) -> np.ndarray:
    """Random-walk chains, shuffled: neighbours k apart along a chain have cos ~ 1/sqrt(1 + k step^2), so a threshold
    between the one-hop and two-hop value makes transitive closure and the reference's greedy one-hop merge differ."""
    rng = np.random.default_rng(seed)
    rows = []
    for _ in range(n_chains):
        x = rng.standard_normal(dim)
        x /= np.linalg.norm(x)
        for _ in range(length):
            rows.append(x.copy())
            n = rng.standard_normal(dim)
            n /= np.linalg.norm(n)
            x = x + step * n
            x /= np.linalg.norm(x)
    x = np.asarray(rows, np.float32)
    return x[rng.permutation(len(x))]


def cluster_cases():
    """(name, rows) of the clustering golden set (tests/golden/make_cluster_golden.py); rows are raw (un-normalised)."""
    a = clustered(41, 60, 5, noise=0.6)                       # clean clusters, cos ~ 0.73 inside
    b = chains(42, 24, 12, 1.2)                               # online threshold 0.45 sits between hop 2 and hop 3
    c = chains(43, 30, 10, 0.6)                               # merge threshold 0.8 sits between hop 1 and hop 2
    d = np.concatenate([clustered(44, 40, 4, noise=0.35), clustered(45, 25, 6, noise=0.2)])   # tight: cos ~ 0.89 / 0.96
    d = d[np.random.default_rng(46).permutation(len(d))]
    scale = np.random.default_rng(47).uniform(3.0, 25.0, (len(d), 1)).astype(np.float32)
    return [("clusters", a), ("chains_wide", b), ("chains_tight", c), ("mixed", d * scale)]
