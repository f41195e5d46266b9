#!/usr/bin/env python
"""bench.py -- headline benchmark of the flink-ms hot paths on B200 (contract: task statement section 4).

    python bench.py --gpus N --steps K --warmup W            # our arm (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on host cores

Primary line (BASELINE.json configs[1]): ALS rank 64, 1M users x 100k items, 100M synthetic ratings; one "step" is
one ALS iteration (users half-step + items half-step) over all ratings.  `value` = ratings/s/iter with the CSR
pair and the factor matrices resident in HBM; `e2e` = the same metric through the C-ABI call fms_als_fit with
host (pinned) COO buffers in and host factor arrays out.  The SVM path (configs[3], 10M x 1024 dense, one epoch
per CoCoA round) is measured in the same run and reported under "svm" in the same JSON line.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def load_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu capture (profiles/)."""
    p = os.path.join(ROOT, "profiles", "r01_traffic.json")
    return json.load(open(p)) if os.path.exists(p) else {}


def full_config(args, world):
    return world == 1 and (args.users, args.items, args.nnz, args.rank) == (1_000_000, 100_000, 100_000_000, 64)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock + throttle reasons of one GPU while the timed region runs (pynvml, 50 ms period)."""
    BITS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
            0x80: "hw_power_brake_slowdown"}

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop, self._t = [], set(), None, threading.Event(), None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # noqa: BLE001
            self.nv = None
            log("clock sampler unavailable:", e)

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.BITS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ workloads
def als_bytes_half(nnz, n_rows, d):
    """Algorithmic bytes of one half-step: nnz*(4 idx + 4 val + 4d row) + N*(4 ptr + 4d out)."""
    return nnz * (8 + 4 * d) + n_rows * (4 + 4 * d)


def svm_bytes_step(dim):
    """Algorithmic bytes of one SDCA step: the row (4D) + y (4) + alpha read+write (8).  (SURVEY's 4D+16 adds an
    optional precomputed x.x; we recompute it from the row already in registers.)"""
    return 4 * dim + 12


def host_threads():
    """Threads for the CPU legs: the cores this process may run on.  Not omp_get_max_threads(): torchrun exports
    OMP_NUM_THREADS=1 to every rank, which would turn the reference arm at N > 1 into a single-threaded run."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return max(1, min(n, 256))


def cpu_als_sample(args, d, threads):
    """Bounded CPU sample of the same workload: same rank, same row-length regime for users (~100 ratings/user)."""
    import oracle
    synth = importlib.import_module("flink-ms_b200.synth")
    U, I, nnz = args.cpu_users, args.items, args.cpu_nnz
    u, i, r = synth.als_ratings(U, I, nnz, seed=1234)
    orc = oracle.OracleALS(u, i, r)
    orc.init(d, 42)
    orc.half_step(0, 0.9, threads)           # users exist before the timed iteration
    return orc, len(u), f"ALS rank {d}, {U} users x {I} items, {len(u)} ratings (same generator), 1 iteration"


def run_reference(args, rank, world):
    """Reference arm: the FlinkML CPU algorithm (oracle port -- the Flink jar cannot run here) on host cores."""
    if rank != 0:
        return
    import oracle
    threads = host_threads()
    d = args.rank
    orc, nnz, sample = cpu_als_sample(args, d, threads)
    for _ in range(args.warmup_ref):
        orc.half_step(0, 0.9, threads); orc.half_step(1, 0.9, threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orc.half_step(0, 0.9, threads)
        orc.half_step(1, 0.9, threads)
    dt = time.perf_counter() - t0
    v = nnz * args.steps / dt
    line = {"impl": "reference", "metric": "als_ratings_per_sec_per_iter", "value": v, "unit": "ratings/s/iter",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup_ref, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world),
            "cpu_baseline": {"value": v, "unit": "ratings/s/iter", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "ratings/s/iter", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {"workload": f"ALS rank={args.rank}, {args.users} users x {args.items} items, {args.nnz} synthetic ratings "
                        f"(BASELINE.json configs[1]); lambda=0.9 seed=42; step = 1 ALS iteration (2 half-steps)",
            "parallelism": (f"rows sharded over {world} GPU(s); "
                            + ("solver warps store each new gather-plane row into every replica over NVLink (CUDA IPC), barrier per half-step"
                               if getattr(args, "als_p2p", 1) else "NCCL all-gather of the gather-plane shards")) if world > 1 else "1 GPU",
            "l2": "inputs larger than L2 (CSR pair 1.6 GB + factors 0.28 GB per iteration vs 126 MB L2); no flush needed"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--users", type=int, default=1_000_000)
    ap.add_argument("--items", type=int, default=100_000)
    ap.add_argument("--nnz", type=int, default=100_000_000)
    ap.add_argument("--rank", type=int, default=64)
    ap.add_argument("--svm-n", type=int, default=10_000_000)
    ap.add_argument("--svm-dim", type=int, default=1024)
    ap.add_argument("--svm-blocks-per-gpu", type=int, default=3552)
    ap.add_argument("--svm-rounds", type=int, default=3)
    ap.add_argument("--skip-svm", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--cpu-users", type=int, default=100_000)
    ap.add_argument("--cpu-nnz", type=int, default=10_000_000)
    ap.add_argument("--warmup-ref", type=int, default=0)
    ap.add_argument("--als-kernel", type=int, default=0)
    ap.add_argument("--als-chunks", type=int, default=1, help="N>1, without --als-p2p 1: solve each shard in this many chunks, all-gather overlapped")
    ap.add_argument("--als-p2p", type=int, default=1, help="N>1: the kernel stores plane rows into all replicas over NVLink (CUDA IPC)")
    ap.add_argument("--svm-kernel", type=int, default=0)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        log("note: --warmup < 3 violates the timing rules; using 3")
        args.warmup = 3

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist

    import flinkms_b200 as fms
    D = importlib.import_module("flink-ms_b200.dist")
    synth = importlib.import_module("flink-ms_b200.synth")
    N = fms._native

    os.environ.setdefault("NCCL_DEBUG", "WARN")           # keep NCCL's version banner off stdout (one JSON line only)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peak, peak_src = load_peaks()
    d = args.rank

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------------------------------------------------------------------------------- ALS inputs (host, pinned)
    t0 = time.perf_counter()
    u_d, i_d, r_d = synth.als_ratings_torch(args.users, args.items, args.nnz, dev, seed=1234)
    nnz = int(u_d.numel())
    h_u = torch.empty(nnz, dtype=torch.int64, pin_memory=True); h_u.copy_(u_d)
    h_i = torch.empty(nnz, dtype=torch.int64, pin_memory=True); h_i.copy_(i_d)
    h_r = torch.empty(nnz, dtype=torch.float64, pin_memory=True); h_r.copy_(r_d)
    del u_d, i_d, r_d
    torch.cuda.empty_cache()
    log(f"[rank {rank}] generated {nnz} ratings in {time.perf_counter() - t0:.1f}s")

    ctx = fms.Context(local_rank)
    ctx.als_set_kernel(args.als_kernel)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    t0 = time.perf_counter()
    n_users, n_items = ctx.als_load(h_u.numpy(), h_i.numpy(), h_r.numpy(), rank, world)
    ctx.als_init(d, 42)
    ctx.synchronize()
    log(f"[rank {rank}] load+index {time.perf_counter() - t0:.2f}s: {n_users} users, {n_items} items")
    backend = D.ContextALSBackend(ctx, dev)
    p2p = bool(world > 1 and args.als_p2p and backend.enable_p2p())
    dals = D.DistributedALS(backend, d, max_chunks=args.als_chunks)

    for _ in range(args.warmup):
        dals.iteration(0.9)
    barrier()
    launches0 = ctx.launch_count
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    e_start, e_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        e_start.record(stream)
        for k in range(args.steps):
            ev[k][0].record(stream)
            dals.half_step(0, 0.9)
            ev[k][1].record(stream)
            dals.half_step(1, 0.9)
            ev[k][2].record(stream)
        e_end.record(stream)
        barrier()
    launches = ctx.launch_count - launches0
    ms_total = max_over_ranks(e_start.elapsed_time(e_end))
    ms_step = ms_total / args.steps
    value = nnz * args.steps / (ms_total * 1e-3)
    ms_user = float(np.mean([ev[k][0].elapsed_time(ev[k][1]) for k in range(args.steps)]))
    ms_item = float(np.mean([ev[k][1].elapsed_time(ev[k][2]) for k in range(args.steps)]))
    # roofline of the dominant kernel (one half-step launch; at N>1 the interval also holds the all-gather)
    shard = 1.0 / world
    by_user = als_bytes_half(nnz, n_users, d) * shard
    by_item = als_bytes_half(nnz, n_items, d) * shard
    dom = "users" if ms_user >= ms_item else "items"
    dom_bytes, dom_ms = (by_user, ms_user) if dom == "users" else (by_item, ms_item)
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    iter_gbs = (by_user + by_item) / (ms_step * 1e-3) / 1e9
    traffic = load_traffic()
    roofline = {"bound": "hbm", "kernel": f"ALS half-step ({dom})", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic.get(f"als_half_step_{dom}") if full_config(args, world) else None,
                "traffic_source": traffic.get("source"), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom_bytes, "ms_per_launch": dom_ms,
                "iteration": {"achieved": iter_gbs, "frac": iter_gbs / peak, "ms_users": ms_user, "ms_items": ms_item}}
    clocks = clk.summary()

    # ---------------------------------------------------------------------------------------- e2e through the C-ABI
    e2e = None
    if not args.skip_e2e:
        T = args.steps
        # result buffers in pinned host memory (what a JVM host gets from a registered direct ByteBuffer)
        o_uid = torch.empty(n_users, dtype=torch.int64, pin_memory=True).numpy()
        o_iid = torch.empty(n_items, dtype=torch.int64, pin_memory=True).numpy()
        o_U = torch.empty((n_users, d), dtype=torch.float64, pin_memory=True).numpy()
        o_I = torch.empty((n_items, d), dtype=torch.float64, pin_memory=True).numpy()
        barrier()
        t0 = time.perf_counter()
        if world == 1:
            ctx.als_fit(h_u.numpy(), h_i.numpy(), h_r.numpy(), d, T, 0.9, 42)          # fms_als_fit: H2D, index, T iters
            ids_u, F_u = ctx.als_get_factors(N.USERS, o_uid, o_U)                         # D2H of the result
            ids_i, F_i = ctx.als_get_factors(N.ITEMS, o_iid, o_I)
            h2d, d2h = nnz * 24, int((n_users + n_items) * (8 + 8 * d))
        else:
            # every rank pulls 1/G of the COO over PCIe, the slices are all-gathered over NVLink, every rank indexes the
            # full set (replicated CSR), solves its row shard and emits its own shard of both factor matrices
            d_u, d_i, d_r = D.sharded_upload((h_u, h_i, h_r), dev)                     # torch's stream; tensors owned by torch
            torch.cuda.current_stream().synchronize()
            ctx.als_load_device(d_u.data_ptr(), d_i.data_ptr(), d_r.data_ptr(), nnz, rank, world)
            del d_u, d_i, d_r
            ctx.als_init(d, 42)
            be2 = D.ContextALSBackend(ctx, dev)
            if args.als_p2p:
                be2.enable_p2p()                                                          # part of the timed call: IPC handle exchange
            D.DistributedALS(be2, d, max_chunks=args.als_chunks).fit(T, 0.9, finalize=False)
            ub, ue, _ = D.shard_range(n_users, world, rank)
            ib, ie, _ = D.shard_range(n_items, world, rank)
            ids_u, F_u = ctx.als_get_factor_rows(N.USERS, ub, ue, o_uid, o_U)
            ids_i, F_i = ctx.als_get_factor_rows(N.ITEMS, ib, ie, o_iid, o_I)
            h2d, d2h = nnz * 24, int((n_users + n_items) * (8 + 8 * d))                   # summed over the ranks
        barrier()
        t_e2e = max_over_ranks(time.perf_counter() - t0)
        e2e = {"value": nnz * T / t_e2e, "unit": "ratings/s/iter", "h2d_bytes_per_step": h2d // T,
               "d2h_bytes_per_step": d2h // T, "seconds": t_e2e, "iterations": T,
               "note": ("one fms_als_fit call (H2D of COO, CSR build, T iterations + final users half-step) + D2H of both "
                        "factor matrices into pinned host buffers; bytes are per call / T") if world == 1 else
                       ("per rank: H2D of 1/G of the pinned host COO, NVLink all-gather of the triples, CSR build, T iterations "
                        "+ final users half-step over the sharded driver, D2H of the rank's own shard of both factor "
                        "matrices into pinned host buffers; bytes are summed over ranks, per call / T")}
        assert np.isfinite(F_u).all() and np.isfinite(F_i).all()
    gpu_launches_als = launches

    # ---------------------------------------------------------------------------------------- predict (configs[4])
    predict = None
    if rank == 0 and world == 1:
        rng = np.random.default_rng(7)
        nq = 1_000_000
        # query and result buffers in pinned host memory (a JVM host's direct ByteBuffers)
        qu = torch.empty(nq, dtype=torch.int64, pin_memory=True).numpy()
        qi = torch.empty(nq, dtype=torch.int64, pin_memory=True).numpy()
        qu[:] = rng.integers(0, args.users, nq, dtype=np.int64)
        qi[:] = rng.integers(0, args.items, nq, dtype=np.int64)
        p_out = torch.empty(nq, dtype=torch.float64, pin_memory=True).numpy()
        p_found = torch.empty(nq, dtype=torch.uint8, pin_memory=True).numpy()
        ctx.als_predict(qu, qi, p_out, p_found)                                      # warm-up (fp64 model copy, scratch)
        t0 = time.perf_counter()
        for _ in range(5):
            out, found = ctx.als_predict(qu, qi, p_out, p_found)
        dtp = (time.perf_counter() - t0) / 5
        predict = {"metric": "als_lookups_per_sec", "value": nq / dtp, "unit": "lookups/s", "batch": nq,
                   "found_fraction": float(found.mean()),
                   "note": "fms_als_predict through the C-ABI: pinned host ids in, H2D, id lookup + fp64 dot on device, D2H into pinned host buffers"}

    # ---------------------------------------------------------------------------------------- CPU baseline (rank 0, N=1)
    cpu = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        import oracle
        threads = host_threads()
        orc, cn, sample = cpu_als_sample(args, d, threads)
        t0 = time.perf_counter()
        orc.half_step(0, 0.9, threads)
        orc.half_step(1, 0.9, threads)
        dt = time.perf_counter() - t0
        cpu = {"value": cn / dt, "unit": "ratings/s/iter", "cores": threads, "kind": "port", "sample": sample,
               "seconds": dt}

    # ---------------------------------------------------------------------------------------- SVM (configs[3])
    svm = None
    ctx.close()
    del h_u, h_i, h_r
    torch.cuda.empty_cache()
    if not args.skip_svm:
        svm = bench_svm(args, fms, D, synth, dev, rank, world, local_rank, peak, peak_src, barrier, max_over_ranks)

    if rank == 0:
        line = {"metric": "als_ratings_per_sec_per_iter", "value": value, "unit": "ratings/s/iter", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(args, world), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": gpu_launches_als, "clocks": clocks, "svm": svm, "predict": predict}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def bench_svm(args, fms, D, synth, dev, rank, world, local_rank, peak, peak_src, barrier, max_over_ranks):
    import torch
    n, dim = args.svm_n, args.svm_dim
    K = args.svm_blocks_per_gpu * world
    t0 = time.perf_counter()
    X, y, _ = synth.svm_data_torch(n, dim, dev, seed=1234)
    torch.cuda.synchronize()
    log(f"[rank {rank}] SVM data {n}x{dim} generated in {time.perf_counter() - t0:.1f}s")
    bid = (np.arange(n, dtype=np.int64) % K).astype(np.int32)
    ctx = fms.Context(local_rank)
    ctx.svm_set_kernel(args.svm_kernel)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    ctx.svm_load(None, y.double().cpu().numpy(), bid, K, rank, world, 1.0, 1.0, 42, dim=dim, device_ptr=X.data_ptr())
    dsvm = D.DistributedSVM(D.ContextSVMBackend(ctx, dev))
    rounds = args.svm_rounds
    for _ in range(3):
        dsvm.round(-1)
    barrier()
    l0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clk:
        e0.record(stream)
        for _ in range(rounds):
            dsvm.round(-1)
        e1.record(stream)
        barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / rounds
    w = ctx.svm_get_weights(dim)
    assert np.isfinite(w).all()
    acc = float(((X[:100000] @ torch.from_numpy(w).float().to(dev)).sign().cpu().numpy() == y[:100000].cpu().numpy()).mean())
    steps_per_rank = n / world
    achieved = svm_bytes_step(dim) * steps_per_rank / (ms * 1e-3) / 1e9
    out = {"metric": "svm_samples_per_sec_per_epoch", "value": n / (ms * 1e-3), "unit": "samples/s/epoch",
           "ms_per_round": ms, "rounds": rounds, "warmup": 3, "dtype": "f32",
           "config": {"workload": f"CoCoA/SDCA linear SVM, {n} x {dim} dense fp32 synthetic (BASELINE.json configs[3]); "
                                  f"K={K} blocks (block_id = i mod K), H = n_k (one epoch per round), lambda=1 "
                                  f"stepsize=1 seed=42", "l2": "X is 41 GB, far larger than L2"},
           "roofline": {"bound": "hbm", "kernel": "SVM block sweep", "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": achieved / peak,
                        "traffic": load_traffic().get("svm_block_sweep") if (world == 1 and n == 10_000_000 and dim == 1024) else None,
                        "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": svm_bytes_step(dim) * steps_per_rank, "ms_per_launch": ms},
           "gpu_launches": ctx.launch_count - l0, "train_accuracy_first_100k": acc, "clocks": clk.summary()}
    ctx.close()
    return out


if __name__ == "__main__":
    main()
