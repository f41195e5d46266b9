"""torchrun --nproc-per-node N tools/dist_check.py : sharded ALS / SVM over NCCL must reproduce the 1-GPU run.
ALS: bitwise equal factors (same per-row arithmetic).  SVM: equal up to the all-reduce summation order."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import flinkms_b200 as fms

D = importlib.import_module("flink-ms_b200.dist")
synth = importlib.import_module("flink-ms_b200.synth")
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.environ.setdefault("NCCL_DEBUG", "WARN")
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
ok = True
# tensor-core path with plain and with chunked / overlapped plane exchange, generic path (fp32 exchange)
for d, chunks in ((64, 1), (64, 4), (64, 0), (16, 1)):          # chunks 0: peer-to-peer plane stores from the kernel
    u, i, r = synth.als_ratings(3001 if chunks == 1 else 4096, 403 if chunks == 1 else 512, 90_000, seed=5)
    ctx = fms.Context(lr)
    ctx.als_load(u, i, r, rank, world)
    ctx.als_init(d, 42)
    be = D.ContextALSBackend(ctx, dev)
    if chunks == 0:
        assert be.enable_p2p(), "peer mapping failed"
    dals = D.DistributedALS(be, d, max_chunks=max(chunks, 1), min_chunk_rows=8)
    dals.fit(3, 0.9)
    if chunks > 1:
        assert dals._overlap_plan(0)[1] > 1, "chunked exchange was not exercised"
    Xs, Ys = ctx.als_get_factors(0)[1], ctx.als_get_factors(1)[1]
    ref = fms.Context(lr)
    ref.als_fit(u, i, r, d, 3, 0.9, 42)
    X1, Y1 = ref.als_get_factors(0)[1], ref.als_get_factors(1)[1]
    same = np.array_equal(Xs, X1) and np.array_equal(Ys, Y1)
    print(f"[rank {rank}] ALS rank {d} chunks {chunks}: sharded == single: {same}  max diff {np.abs(Xs - X1).max():.3e}", flush=True)
    ok &= same
    ctx.close(); ref.close()
n, dim, K = 20000, 256, 24
X, y, _ = synth.svm_data(n, dim, seed=6)
bid = (np.arange(n) % K).astype(np.int32)
ctx = fms.Context(lr)
ctx.svm_load(X, y, bid, K, rank, world, 1.0, 1.0, 42)
D.DistributedSVM(D.ContextSVMBackend(ctx, dev)).fit(4, -1)
ws = ctx.svm_get_weights(dim)
ref = fms.Context(lr)
w1 = ref.svm_fit(X, y, bid, K, 4, -1, 1.0, 1.0, 42)
err = np.abs(ws - w1).max() / np.abs(w1).max()
print(f"[rank {rank}] SVM sharded (NCCL all-reduce) vs single rel err {err:.3e}", flush=True)
ok &= err < 1e-5
# the same fit with the fused exchange kernel over IPC-mapped buffers; every rank must end with the SAME bits
ctx2 = fms.Context(lr)
ctx2.svm_load(X, y, bid, K, rank, world, 1.0, 1.0, 42)
be = D.ContextSVMBackend(ctx2, dev)
be.enable_p2p()
D.DistributedSVM(be).fit(4, -1)
ctx2.svm_exchange_status()
wx = ctx2.svm_get_weights(dim)
errx = np.abs(wx - w1).max() / np.abs(w1).max()
wt = torch.from_numpy(wx).to(dev)
lo, hi = wt.clone(), wt.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
same_bits = bool((lo == hi).all().item())
print(f"[rank {rank}] SVM sharded (exchange kernel) vs single rel err {errx:.3e}; identical on all ranks: {same_bits}", flush=True)
ok &= errx < 1e-5 and same_bits
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
