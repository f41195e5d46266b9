"""Phase timing of the e2e path (GPU box): H2D + CSR build (fms_als_load), init, T iterations, D2H read-out."""
import importlib
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import flinkms_b200 as fms

synth = importlib.import_module("flink-ms_b200.synth")
dev = torch.device("cuda", 0)
U, I, nnz = 1_000_000, 100_000, 100_000_000
u_d, i_d, r_d = synth.als_ratings_torch(U, I, nnz, dev, seed=1234)
n = int(u_d.numel())
h_u = torch.empty(n, dtype=torch.int64, pin_memory=True); h_u.copy_(u_d)
h_i = torch.empty(n, dtype=torch.int64, pin_memory=True); h_i.copy_(i_d)
h_r = torch.empty(n, dtype=torch.float64, pin_memory=True); h_r.copy_(r_d)
del u_d, i_d, r_d
torch.cuda.empty_cache()
ctx = fms.Context(0)
for rep in range(2):
    t0 = time.perf_counter(); ctx.als_load(h_u.numpy(), h_i.numpy(), h_r.numpy()); ctx.synchronize(); t1 = time.perf_counter()
    ctx.als_init(64, 42); ctx.synchronize(); t2 = time.perf_counter()
    for _ in range(5):
        ctx.als_half_step(0, 0.9); ctx.als_half_step(1, 0.9)
    ctx.als_half_step(0, 0.9); ctx.synchronize(); t3 = time.perf_counter()
    ids_u, F_u = ctx.als_get_factors(0); ids_i, F_i = ctx.als_get_factors(1); t4 = time.perf_counter()
    print(f"rep {rep}: load {t1 - t0:.3f}s init {t2 - t1:.3f}s iters {t3 - t2:.3f}s readout {t4 - t3:.3f}s total {t4 - t0:.3f}s", flush=True)
