"""Seeded synthetic inputs of the shapes BASELINE.md section 5 names (the reference's own model-generator/
uses an unseeded scala.util.Random and is "not for quality" by its own header, ALSModelGenerator.scala:28-32).

numpy generators serve the parity tests; the torch generators build bench-scale inputs on the GPU.
"""
from __future__ import annotations

import numpy as np


def als_ratings(n_users: int, n_items: int, nnz: int, seed: int = 1234, id_stride=(1, 1), id_offset=(0, 0)):
    """Distinct uniform (u, i) pairs; rating = clip(round(N(3.5, 1) * 2) / 2, 0.5, 5).  ids = dense * stride + offset."""
    rng = np.random.default_rng(seed)
    keys = np.unique(rng.integers(0, n_users * n_items, size=int(nnz * 1.05) + 16, dtype=np.int64))
    rng.shuffle(keys)
    keys = keys[:nnz]
    u = keys // n_items
    i = keys % n_items
    r = np.clip(np.round(rng.normal(3.5, 1.0, size=len(keys)) * 2.0) / 2.0, 0.5, 5.0)
    return (u * id_stride[0] + id_offset[0]).astype(np.int64), (i * id_stride[1] + id_offset[1]).astype(np.int64), r


def svm_data(n: int, dim: int, seed: int = 1234, noise: float = 0.1):
    """X ~ N(0,1) fp32, w* ~ N(0,1), y = sign(X w* + noise * N(0,1)) in {+1,-1}."""
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, dim), dtype=np.float32)
    w_star = rng.standard_normal(dim)
    y = np.sign(X.astype(np.float64) @ w_star + noise * rng.standard_normal(n))
    y[y == 0] = 1.0
    return X, y, w_star


def svm_sparse(n: int, dim: int, density: float = 0.2, seed: int = 1234):
    """Sparse rows as CSR (ptr, idx, val) with fp32-representable values (LibSVM-style input)."""
    rng = np.random.default_rng(seed)
    ptr = [0]
    idx, val = [], []
    for _ in range(n):
        k = max(1, int(rng.binomial(dim, density)))
        cols = np.sort(rng.choice(dim, size=k, replace=False))
        idx.append(cols)
        val.append(rng.standard_normal(k).astype(np.float32).astype(np.float64))
        ptr.append(ptr[-1] + k)
    idx = np.concatenate(idx).astype(np.int32)
    val = np.concatenate(val)
    w_star = rng.standard_normal(dim)
    ptr = np.asarray(ptr, np.int64)
    raw = np.array([val[ptr[r]:ptr[r + 1]] @ w_star[idx[ptr[r]:ptr[r + 1]]] for r in range(n)])
    y = np.sign(raw + 0.1 * rng.standard_normal(n))
    y[y == 0] = 1.0
    return (ptr, idx, val), y


# ------------------------------------------------------------------------------------------ bench scale (torch)
def als_ratings_torch(n_users: int, n_items: int, nnz: int, device, seed: int = 1234):
    """Same distribution, generated on `device`; returns int64 user, int64 item, float64 rating tensors."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    keys = torch.randint(0, n_users * n_items, (int(nnz * 1.02) + 1024,), generator=g, device=device, dtype=torch.int64)
    keys = torch.unique(keys)
    perm = torch.randperm(keys.numel(), generator=g, device=device)[:nnz]
    keys = keys[perm]
    u = keys // n_items
    i = keys % n_items
    r = torch.clamp(torch.round(torch.normal(3.5, 1.0, (keys.numel(),), generator=g, device=device) * 2.0) / 2.0, 0.5, 5.0)
    return u, i, r.double()


def svm_data_torch(n: int, dim: int, device, seed: int = 1234, noise: float = 0.1, chunk: int = 1 << 20):
    """X (n x dim fp32, on device), y (float64 on host as the ABI wants), generated chunk by chunk."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    w_star = torch.randn(dim, generator=g, device=device)
    X = torch.empty((n, dim), dtype=torch.float32, device=device)
    y = torch.empty(n, dtype=torch.float32, device=device)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        X[s:e].normal_(generator=g)
        raw = X[s:e] @ w_star + noise * torch.randn(e - s, generator=g, device=device)
        y[s:e] = torch.where(raw >= 0, 1.0, -1.0)
    return X, y, w_star
