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
    for name in ("r02_traffic.json", "r01_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            return json.load(open(p))
    return {}


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


def cgroup_cpu_quota():
    """CPUs the container may use per its cgroup CPU bandwidth limit (None = unlimited / unknown)."""
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]                      # cgroup v2
        return None if q == "max" else float(q) / float(per)
    except Exception:  # noqa: BLE001
        pass
    try:
        q = float(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())                   # cgroup v1
        per = float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
        return None if q <= 0 else q / per
    except Exception:  # noqa: BLE001
        return None


def host_threads():
    """Threads for the CPU legs: the cores this process may run on, capped by the cgroup CPU quota (more threads than
    quota only buys throttling: round 1 saw a 6x swing of the CPU arm between two "128-core" boxes).  Not
    omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 to every rank, which would turn the reference arm at
    N > 1 into a single-threaded run."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    q = cgroup_cpu_quota()
    if q is not None:
        n = min(n, max(1, int(q)))
    return max(1, min(n, 256))


def cpu_info(threads):
    try:
        aff = len(os.sched_getaffinity(0))
    except AttributeError:
        aff = os.cpu_count()
    return {"cores": threads, "affinity_cores": aff, "cgroup_cpu_quota": cgroup_cpu_quota()}


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


def cpu_svm_sample(args, threads, rounds=10):
    """Bounded CPU sample of the SVM workload: same dim, same block size n_k as the GPU run (K scaled with n), one epoch
    per round, `rounds` rounds of the oracle's CoCoA loop on all host threads (blocks are independent inside a round)."""
    import oracle
    synth = importlib.import_module("flink-ms_b200.synth")
    nk = max(1, args.svm_n // args.svm_blocks)
    Ks = max(1, args.cpu_svm_n // nk)
    ns = Ks * nk
    X, y, _ = synth.svm_data(ns, args.svm_dim, seed=1234)
    bid = (np.arange(ns, dtype=np.int64) % Ks).astype(np.int32)
    t0 = time.perf_counter()
    oracle.svm_fit(X, y, bid, Ks, rounds, -1, 1.0, 1.0, 42, 0, threads)
    dt = time.perf_counter() - t0
    info = cpu_info(threads)
    info.update({"value": ns * rounds / dt, "unit": "samples/s/epoch", "kind": "port", "seconds": dt,
                 "sample": f"CoCoA/SDCA {ns} x {args.svm_dim} dense fp32 (same generator), K={Ks} blocks of n_k={nk} rows, "
                           f"H = n_k, {rounds} rounds, fp64 oracle port, {threads} threads"})
    return info


def run_reference(args, rank, world):
    """Reference arm: the FlinkML CPU algorithm (oracle port -- the Flink jar cannot run here) on host cores, on a
    BOUNDED SAMPLE of the workload (stated in config.workload and cpu_baseline.sample)."""
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
    cfg = workload_config(args, world, None)
    cfg["workload"] += f" -- CPU arm timed on a bounded sample of it: {sample}"
    cfg["parallelism"] = f"OpenMP over rows, {threads} host threads"
    cb = cpu_info(threads)
    cb.update({"value": v, "unit": "ratings/s/iter", "kind": "port", "sample": sample})
    svm = None
    if not args.skip_svm:
        sv = cpu_svm_sample(args, threads)
        svm = {"impl": "reference", "metric": "svm_samples_per_sec_per_epoch", "value": sv["value"], "unit": "samples/s/epoch",
               "cpu_baseline": sv, "e2e": {"value": sv["value"], "unit": "samples/s/epoch", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    line = {"impl": "reference", "metric": "als_ratings_per_sec_per_iter", "value": v, "unit": "ratings/s/iter",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup_ref, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": cfg, "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "ratings/s/iter", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "svm": svm}
    print(json.dumps(line), flush=True)


def workload_config(args, world, p2p):
    """p2p: what the run really did (None: not applicable), not what the flags asked for."""
    cfgno = {64: "configs[1]", 128: "configs[2]"}.get(args.rank, "configs[1] shape")
    if world == 1:
        par = "1 GPU"
    elif p2p:
        par = (f"rows sharded over {world} GPUs, one process per GPU; solver warps store each new gather-plane row into every "
               f"replica over NVLink (CUDA IPC peer mappings), the half-step kernels order themselves with epoch flags in the replicas (release/acquire at system scope) -- no host collective between half-steps")
    else:
        par = f"rows sharded over {world} GPUs, one process per GPU; NCCL all-gather of the solved shard after every half-step"
    return {"workload": f"ALS rank={args.rank}, {args.users} users x {args.items} items, {args.nnz} synthetic ratings "
                        f"(BASELINE.json {cfgno}); lambda=0.9 seed=42; step = 1 ALS iteration (2 half-steps)",
            "parallelism": par,
            "l2": "inputs larger than L2 (CSR pair 1.6 GB + factors 0.28 GB per iteration vs 126 MB L2); no flush needed"}


# ------------------------------------------------------------------------------------------------ bench-scale parity
def _partner_rows_f64(backend, partner, d, idx_t):
    """The exact partner values the half-step kernel consumed for dense partner indices idx_t (device int64 tensor):
    hi + lo / 2048 of the fp16 gather planes on the tensor-core path, the fp32 rows otherwise.  float64, on the host."""
    planes = backend.plane_tensors(partner, d)
    if planes is not None:
        pl = planes[0].index_select(0, idx_t)
        return (pl[:, :d].double() + pl[:, d:].double() / 2048.0).cpu().numpy()
    return backend.factor_tensor(partner, d).index_select(0, idx_t).double().cpu().numpy()


def als_parity_side(ctx, backend, which, lam, d, h_t, h_p, h_r, n_side, n_partner, rank, world, dev, samples, shard_range):
    """Rows of side `which` (0 users / 1 items) were just solved from the partner side's current values.  Take `samples`
    random rows of this rank's shard, rebuild each row's normal equations in fp64 from the exact partner values the
    kernel read and the host COO, solve them with LAPACK and compare with the device's row (max-norm relative error)."""
    import torch
    rb, re_, _ = shard_range(n_side, world, rank)
    if re_ <= rb:
        return 0.0, 0
    rng = np.random.default_rng(4242 + 17 * which + rank)
    rows = np.sort(rng.choice(re_ - rb, size=min(samples, re_ - rb), replace=False)) + rb
    lut = np.full(n_side, -1, np.int32)
    lut[rows] = np.arange(len(rows), dtype=np.int32)
    slot = lut[h_t]                                                 # bench ids are dense (checked by the caller)
    sel = np.flatnonzero(slot >= 0)
    order = np.argsort(slot[sel], kind="stable")
    sel = sel[order]
    starts = np.searchsorted(slot[sel], np.arange(len(rows) + 1))
    P = _partner_rows_f64(backend, 1 - which, d, torch.from_numpy(h_p[sel]).to(dev))
    R = h_r[sel]
    worst = 0.0
    for k, row in enumerate(rows):
        a, b = starts[k], starts[k + 1]
        Y = P[a:b]
        A = Y.T @ Y + lam * (b - a) * np.eye(d)
        x_ref = np.linalg.solve(A, Y.T @ R[a:b])
        _, x_dev = ctx.als_get_factor_rows(which, int(row), int(row) + 1)
        worst = max(worst, float(np.abs(x_dev[0] - x_ref).max() / np.abs(x_ref).max()))
    return worst, len(rows)


def svm_parity(ctx, dsvm, oracle, X, y_host, n, dim, K, rank, world, rounds_done, blocks_to_check=8, seed=42, lam=1.0):
    """Replays `blocks_to_check` whole blocks of ONE more CoCoA round in fp64 (numpy restatement of localDualMethod /
    maximize, the java.util.Random index stream from the oracle) from the device's w / alpha snapshot and compares the
    device's alpha after the round and the block's delta-w."""
    w0 = ctx.svm_get_weights(dim)
    a0 = ctx.svm_get_alpha()
    dsvm.round(-1)
    a1 = ctx.svm_get_alpha()
    ln = lam * n
    worst_a, worst_dw, checked = 0.0, 0.0, 0
    for k in range(rank, K, world):
        if checked >= blocks_to_check:
            break
        rows = np.arange(k, n, K)
        nk = len(rows)
        if nk == 0:
            continue
        stream = oracle.svm_index_stream(seed, k, nk, (rounds_done + 1) * nk)[-nk:]
        import torch
        Xk = X[torch.from_numpy(rows).to(X.device)].double().cpu().numpy()
        yk = y_host[rows]
        wl, tmp, dw = w0.copy(), a0[rows].copy(), np.zeros(dim)
        for idx in stream:
            x, yy = Xk[idx], yk[idx]
            g = (yy * float(x @ wl) - 1.0) * ln
            if g != 0.0:
                q = float(x @ x)
                an = min(max(tmp[idx] - g / q, 0.0), 1.0) if q != 0.0 else 1.0
                step = x * (yy * (an - tmp[idx]) / ln)
                wl += step; dw += step
                tmp[idx] = an
        worst_a = max(worst_a, float(np.abs(a1[rows] - tmp).max()))
        dw_dev = ((a1[rows] - a0[rows]) * yk / ln) @ Xk
        worst_dw = max(worst_dw, float(np.abs(dw_dev - dw).max() / max(np.abs(dw).max(), 1e-300)))
        checked += 1
    return {"blocks_checked": checked, "alpha_max_abs_diff": worst_a, "delta_w_max_rel": worst_dw}


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
    ap.add_argument("--svm-blocks", type=int, default=28416,
                    help="CoCoA blocks K, the SAME for every N (the model depends on K): 3552 warps per GPU at 8 GPUs")
    ap.add_argument("--svm-rounds", type=int, default=10)
    ap.add_argument("--cpu-svm-n", type=int, default=500_000)
    ap.add_argument("--skip-parity", action="store_true")
    ap.add_argument("--parity-rows", type=int, default=512)
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
    ap.add_argument("--svm-p2p", type=int, default=1, help="N>1: fused push / flag / sum / apply kernel over IPC-mapped buffers instead of NCCL")
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

    # ---------------------------------------------------------------------------------------- parity at bench scale
    # Not timed.  The device state here is the one the timed loop left: items were just solved from the current users.
    parity = None
    if not args.skip_parity:
        import zlib
        t0 = time.perf_counter()
        hu, hi, hr = h_u.numpy(), h_i.numpy(), h_r.numpy()
        dense_ids = (n_users == int(hu.max()) + 1) and (n_items == int(hi.max()) + 1)      # synthetic ids are 0 .. N-1
        parity = {"tolerance": 1e-5, "reference": "fp64 LAPACK solve of each sampled row's normal equations from the exact partner "
                                                  "values the kernel read (hi + lo/2048 of the gather planes) and the host COO"}
        if dense_ids:
            e_items, n_i = als_parity_side(ctx, backend, 1, 0.9, d, hi, hu, hr, n_items, n_users, rank, world, dev, args.parity_rows, D.shard_range)
            dals.half_step(0, 0.9)                                                       # users from the current items
            ctx.synchronize()
            barrier()
            e_users, n_u = als_parity_side(ctx, backend, 0, 0.9, d, hu, hi, hr, n_users, n_items, rank, world, dev, args.parity_rows, D.shard_range)
            parity.update({"als_items_max_rel": max_over_ranks(e_items), "als_users_max_rel": max_over_ranks(e_users),
                           "rows_checked_per_rank": [n_u, n_i]})
            parity["max_rel"] = max(parity["als_items_max_rel"], parity["als_users_max_rel"])
        else:
            parity["skipped"] = "ids are not dense 0..N-1"
        # CRC of both fp32 factor matrices after (warmup + steps) iterations + 1 users half-step: per-row arithmetic does
        # not depend on the sharding, so the value must be IDENTICAL at every N
        dals.finalize()
        ctx.synchronize()
        barrier()
        if rank == 0:
            fu = backend.factor_tensor(0, d)[:n_users].cpu().numpy()
            fi = backend.factor_tensor(1, d)[:n_items].cpu().numpy()
            parity["factors_crc"] = f"{zlib.crc32(fu.tobytes()):08x}-{zlib.crc32(fi.tobytes()):08x}"
            parity["factors_crc_of"] = f"fp32 U ({n_users}x{d}) and I ({n_items}x{d}) after {args.warmup + args.steps} iterations + 1 users half-step"
            parity["bad_rows_last_half_step"] = ctx.als_last_info()
        parity["seconds"] = time.perf_counter() - t0

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
    if True:
        # every rank holds a full replica of the model after finalize(): the queries are sharded round-robin over the ranks
        # ("predict: replicas, shard queries"), time = max over ranks, value = all queries / that time
        D.DistributedALS(D.ContextALSBackend(ctx, dev), d).refresh_replicas()       # the e2e leg skipped the final all-gather
        rng = np.random.default_rng(7)
        nq_total = 1_000_000
        all_u = rng.integers(0, args.users, nq_total, dtype=np.int64)
        all_i = rng.integers(0, args.items, nq_total, dtype=np.int64)
        nq = len(all_u[rank::world])
        # query and result buffers in pinned host memory (a JVM host's direct ByteBuffers)
        qu = torch.empty(nq, dtype=torch.int64, pin_memory=True).numpy()
        qi = torch.empty(nq, dtype=torch.int64, pin_memory=True).numpy()
        qu[:] = all_u[rank::world]
        qi[:] = all_i[rank::world]
        p_out = torch.empty(nq, dtype=torch.float64, pin_memory=True).numpy()
        p_found = torch.empty(nq, dtype=torch.uint8, pin_memory=True).numpy()
        ctx.als_predict(qu, qi, p_out, p_found)                                      # warm-up (fp64 model copy, scratch)
        barrier()
        t0 = time.perf_counter()
        for _ in range(5):
            out, found = ctx.als_predict(qu, qi, p_out, p_found)
        dtp = max_over_ranks((time.perf_counter() - t0) / 5)
        predict = {"metric": "als_lookups_per_sec", "value": nq_total / dtp, "unit": "lookups/s", "batch": nq_total,
                   "found_fraction": float(found.mean()), "replicas": world,
                   "note": "fms_als_predict through the C-ABI: pinned host ids in, H2D, id lookup + fp64 dot on device, D2H into pinned "
                           "host buffers; one model replica per rank, queries sharded round-robin, max over ranks"}

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
        cpu = cpu_info(threads)
        cpu.update({"value": cn / dt, "unit": "ratings/s/iter", "kind": "port", "sample": sample, "seconds": dt})

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
                "config": workload_config(args, world, p2p), "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
                "gpu_launches": gpu_launches_als, "clocks": clocks, "parity": parity, "svm": svm, "predict": predict}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def bench_svm(args, fms, D, synth, dev, rank, world, local_rank, peak, peak_src, barrier, max_over_ranks):
    import torch
    n, dim = args.svm_n, args.svm_dim
    K = args.svm_blocks                                   # fixed for every N: the trained model depends on K
    t0 = time.perf_counter()
    X, y, _ = synth.svm_data_torch(n, dim, dev, seed=1234)
    torch.cuda.synchronize()
    log(f"[rank {rank}] SVM data {n}x{dim} generated in {time.perf_counter() - t0:.1f}s")
    bid = (np.arange(n, dtype=np.int64) % K).astype(np.int32)
    y_host = y.double().cpu().numpy()
    ctx = fms.Context(local_rank)
    ctx.svm_set_kernel(args.svm_kernel)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    ctx.svm_load(None, y_host, bid, K, rank, world, 1.0, 1.0, 42, dim=dim, device_ptr=X.data_ptr())
    svm_backend = D.ContextSVMBackend(ctx, dev)
    svm_p2p = bool(world > 1 and args.svm_p2p and svm_backend.enable_p2p())
    dsvm = D.DistributedSVM(svm_backend)
    rounds, warm = args.svm_rounds, 3
    for _ in range(warm):
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
    launches = ctx.launch_count - l0
    ms = max_over_ranks(e0.elapsed_time(e1)) / rounds
    w = ctx.svm_get_weights(dim)
    assert np.isfinite(w).all()
    acc = float(((X[:100000] @ torch.from_numpy(w).float().to(dev)).sign().cpu().numpy() == y[:100000].cpu().numpy()).mean())
    steps_per_rank = n / world
    achieved = svm_bytes_step(dim) * steps_per_rank / (ms * 1e-3) / 1e9
    out = {"metric": "svm_samples_per_sec_per_epoch", "value": n / (ms * 1e-3), "unit": "samples/s/epoch",
           "ms_per_round": ms, "rounds": rounds, "warmup": warm, "dtype": "f32", "n_gpus": world,
           "config": {"workload": f"CoCoA/SDCA linear SVM, {n} x {dim} dense fp32 synthetic (BASELINE.json configs[3]); "
                                  f"K={K} blocks for every N (block_id = i mod K, n_k = {n // K}), H = n_k (one epoch per round), "
                                  f"lambda=1 stepsize=1 seed=42",
                      "parallelism": "1 GPU" if world == 1 else (f"blocks k = g mod {world} on GPU g; " + (
                          "delta-w pushed into every rank's exchange buffer over NVLink (CUDA IPC), epoch flags, rank-ordered sum + apply "
                          "in ONE kernel per round -- no collective call" if svm_p2p else "delta-w all-reduced per round (NCCL)")),
                      "l2": "X is 41 GB, far larger than L2"},
           "roofline": {"bound": "hbm", "kernel": "SVM block sweep", "achieved": achieved, "peak": peak, "unit": "GB/s",
                        "frac": achieved / peak,
                        "traffic": load_traffic().get("svm_block_sweep") if (world == 1 and n == 10_000_000 and dim == 1024) else None,
                        "peak_source": peak_src,
                        "algorithmic_bytes_per_launch": svm_bytes_step(dim) * steps_per_rank, "ms_per_launch": ms},
           "gpu_launches": launches, "train_accuracy_first_100k": acc, "clocks": clk.summary(),
           "w_checksum": {"l2": float(np.sqrt((w * w).sum())), "sum": float(w.sum()), "first4": [float(v) for v in w[:4]],
                          "after_rounds": warm + rounds}}
    # ---- parity at bench scale: replay whole blocks of one more round in fp64 from the device's snapshot
    if not args.skip_parity:
        import oracle
        t0 = time.perf_counter()
        pr = svm_parity(ctx, dsvm, oracle, X, y_host, n, dim, K, rank, world, warm + rounds)
        pr["alpha_max_abs_diff"] = max_over_ranks(pr["alpha_max_abs_diff"])
        pr["delta_w_max_rel"] = max_over_ranks(pr["delta_w_max_rel"])
        pr["blocks_checked"] = f"{pr['blocks_checked']} per rank"
        pr["tolerance"] = "alpha 1e-4 abs, delta-w 1e-4 rel (fp32 device vs fp64 replay of the same steps)"
        pr["seconds"] = time.perf_counter() - t0
        out["parity"] = pr
    ctx.close()
    # ---- e2e through the C-ABI: host fp32 X (pinned) -> fms_svm_fit -> w on the host.  N = 1 only: the sharded run keeps
    #      X on the device and has no host-buffer entry point that avoids uploading the full matrix to every rank.
    out["e2e"] = None
    if world == 1 and not args.skip_e2e:
        try:
            t0 = time.perf_counter()
            hX = torch.empty((n, dim), dtype=torch.float32, pin_memory=True)
            hX.copy_(X)
            torch.cuda.synchronize()
            t_pin = time.perf_counter() - t0
            T = min(rounds, 10)
            ctx2 = fms.Context(local_rank)
            ctx2.svm_set_kernel(args.svm_kernel)
            t0 = time.perf_counter()
            w2 = ctx2.svm_fit(hX.numpy(), y_host, bid, K, T, -1, 1.0, 1.0, 42)
            t_fit = time.perf_counter() - t0
            assert np.isfinite(w2).all()
            out["e2e"] = {"value": n * T / t_fit, "unit": "samples/s/epoch", "seconds": t_fit, "rounds": T,
                          "h2d_bytes_per_step": int((n * dim * 4 + n * 8 + n * 4) // T), "d2h_bytes_per_step": int(dim * 8 // T),
                          "gpu_launches": ctx2.launch_count,
                          "note": f"one fms_svm_fit call: pinned host X ({n * dim * 4 / 1e9:.1f} GB fp32) + labels + block ids H2D, {T} CoCoA rounds, "
                                  f"w D2H; bytes are per call / rounds (pinning + filling the host buffer took {t_pin:.1f} s, not timed)"}
            ctx2.close()
            del hX
        except Exception as e:  # noqa: BLE001  (host memory for 41 GB of pinned X not available)
            out["e2e"] = {"value": None, "error": repr(e)}
    # ---- CPU baseline (rank 0, N = 1): the oracle port on a bounded sample, all host threads
    out["cpu_baseline"] = None
    if rank == 0 and world == 1 and not args.skip_cpu:
        del X
        torch.cuda.empty_cache()
        out["cpu_baseline"] = cpu_svm_sample(args, host_threads())
    return out


if __name__ == "__main__":
    main()
