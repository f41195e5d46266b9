"""Stand-alone SVM throughput probe (GPU box): python tests/svm_probe.py [n] [K ...] -- prints ms/round and GB/s."""
import importlib
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import flinkms_b200 as fms

synth = importlib.import_module("flink-ms_b200.synth")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
Ks = [int(a) for a in sys.argv[2:]] or [1184]
dim = 1024
dev = torch.device("cuda", 0)
X, y, _ = synth.svm_data_torch(n, dim, dev, seed=1234)
yh = y.double().cpu().numpy()
for K in Ks:
    bid = (np.arange(n, dtype=np.int64) % K).astype(np.int32)
    ctx = fms.Context(0)
    ctx.svm_load(None, yh, bid, K, 0, 1, 1.0, 1.0, 42, dim=dim, device_ptr=X.data_ptr())
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    for _ in range(3):
        ctx.svm_sweep(-1); ctx.svm_apply_delta()
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(3):
        ctx.svm_sweep(-1); ctx.svm_apply_delta()
    e1.record(stream)
    ctx.synchronize()
    ms = e0.elapsed_time(e1) / 3
    w = ctx.svm_get_weights(dim)
    acc = float(((X[:100000] @ torch.from_numpy(w).float().to(dev)).sign().cpu().numpy() == y[:100000].cpu().numpy()).mean())
    print(f"K={K} ms/round={ms:.3f} GB/s={(4 * dim + 12) * n / ms / 1e6:.0f} acc={acc:.4f}", flush=True)
    ctx.close()
