"""Stand-alone probe of the tensor-core ALS kernel (run on the GPU box): Gramian / rhs accuracy against fp64 numpy,
then one half-step against the generic kernel.  Prints numbers; used while developing, not a pytest file."""
import importlib
import sys

import numpy as np

sys.path.insert(0, ".")
import flinkms_b200 as fms

synth = importlib.import_module("flink-ms_b200.synth")
N = fms._native
U, I, nnz = (int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (300, 40, 6000)))
u, i, r = synth.als_ratings(U, I, nnz, seed=3)
ctx = fms.Context(0)
ctx.als_load(u, i, r)
ctx.als_init(64, 42)
_, Y = ctx.als_get_factors(N.ITEMS)
uid = np.unique(u)
iid = np.unique(i)
nrows = min(len(uid), 64)
G, b = ctx.debug_als_gramian(N.USERS, 0.9, 0, nrows)
worst_g = worst_b = 0.0
for row in range(nrows):
    sel = u == uid[row]
    V = Y[np.searchsorted(iid, i[sel])]
    Gr = V.T @ V + 0.9 * sel.sum() * np.eye(64)
    br = V.T @ r[sel]
    worst_g = max(worst_g, np.abs(G[row] - Gr).max() / np.abs(Gr).max())
    worst_b = max(worst_b, np.abs(b[row] - br).max() / np.abs(br).max())
print("gramian rel err", worst_g, "rhs rel err", worst_b, flush=True)
ctx.als_init(64, 42)
ctx.als_set_kernel(N.ALS_KERNEL_GENERIC)
ctx.als_half_step(N.USERS, 0.9)
_, Xg = ctx.als_get_factors(N.USERS)
ctx.als_set_kernel(N.ALS_KERNEL_TC)
ctx.als_half_step(N.USERS, 0.9)
_, Xt = ctx.als_get_factors(N.USERS)
print("users tc vs generic", np.abs(Xt - Xg).max() / np.abs(Xg).max(), "bad", ctx.als_last_info(), flush=True)
ctx.als_half_step(N.ITEMS, 0.9)
_, Yt = ctx.als_get_factors(N.ITEMS)
ctx.als_set_kernel(N.ALS_KERNEL_GENERIC)
ctx.als_half_step(N.ITEMS, 0.9)
_, Yg = ctx.als_get_factors(N.ITEMS)
print("items tc vs generic", np.abs(Yt - Yg).max() / np.abs(Yg).max(), flush=True)
