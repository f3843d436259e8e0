"""Seeded synthetic datasets and exact ground truth for the HNSW hot path.

Primary: low-rank latent Gaussian, x = z P + sigma * eps, z ~ N(0, I_r), P ~ N(0, 1/r)^{r x dim}
(r=16, sigma=0.05 for 128-d; r=32 for 768-d).  Secondary: uniform U[0,1)^dim.
Host-side numpy only: the same arrays feed the device index and the CPU oracle.
"""
import math

import numpy as np


def lowrank(n, dim, r=16, sigma=0.05, seed=123, n_queries=0):
    """Returns (data [n,dim] f32, queries [n_queries,dim] f32) drawn from the same distribution."""
    rng = np.random.default_rng(seed)
    P = (rng.standard_normal((r, dim)) / math.sqrt(r)).astype(np.float32)

    def draw(k):
        out = np.empty((k, dim), dtype=np.float32)
        step = 1 << 18
        for s in range(0, k, step):
            e = min(k, s + step)
            z = rng.standard_normal((e - s, r), dtype=np.float32)
            out[s:e] = z @ P + np.float32(sigma) * rng.standard_normal((e - s, dim), dtype=np.float32)
        return out

    x = draw(n)
    q = draw(n_queries) if n_queries else np.empty((0, dim), np.float32)
    return x, q


def uniform(n, dim, seed=123, n_queries=0):
    rng = np.random.default_rng(seed)
    x = rng.random((n, dim), dtype=np.float32)
    q = rng.random((n_queries, dim), dtype=np.float32) if n_queries else np.empty((0, dim), np.float32)
    return x, q


def draw_levels(n, m, seed=42):
    """Injected per-node levels, floor(-ln(u) / ln(m)) with u ~ U[0,1) f64 (reference core.rs:601-605).
    The first node of an empty index ignores its entry (core.rs:393-405)."""
    rng = np.random.default_rng(seed)
    u = rng.random(n)
    u = np.where(u == 0.0, np.nextafter(0.0, 1.0), u)
    return np.floor(-np.log(u) * (1.0 / math.log(m))).astype(np.int32)


def brute_force_topk(x, q, k, device=None, block=4096):
    """Exact top-k by -||x - q||^2 (ties irrelevant for recall).  Uses torch on `device` if given."""
    if device is not None:
        import torch

        xt = torch.from_numpy(x).to(device)
        xn = (xt * xt).sum(1)
        out = np.empty((q.shape[0], k), dtype=np.int64)
        block = max(16, min(block, (1 << 31) // max(1, x.shape[0])))   # the [block, n] distance matrix stays under 8 GiB
        for s in range(0, q.shape[0], block):
            qt = torch.from_numpy(q[s:s + block]).to(device)
            d = xn[None, :] - 2.0 * (qt @ xt.T)
            out[s:s + block] = d.topk(k, dim=1, largest=False).indices.cpu().numpy()
        return out
    xn = (x.astype(np.float64) ** 2).sum(1)
    out = np.empty((q.shape[0], k), dtype=np.int64)
    for s in range(0, q.shape[0], 256):
        d = xn[None, :] - 2.0 * (q[s:s + 256].astype(np.float64) @ x.T.astype(np.float64))
        idx = np.argpartition(d, k - 1, axis=1)[:, :k]
        dd = np.take_along_axis(d, idx, 1)
        out[s:s + 256] = np.take_along_axis(idx, np.argsort(dd, axis=1), 1)
    return out


def recall_at_k(found_ids, gt_ids):
    """Mean |found ∩ gt| / k over queries."""
    k = gt_ids.shape[1]
    hit = 0
    for f, g in zip(found_ids, gt_ids):
        hit += len(set(int(v) for v in f[:k]) & set(int(v) for v in g))
    return hit / (k * len(gt_ids))
