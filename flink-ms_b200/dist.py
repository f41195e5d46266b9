"""One-process-per-GPU sharding of the two hot paths over torch.distributed (NCCL on GPUs, gloo in CPU tests).

ALS: every rank holds the CSR rows of its own shard and full factor replicas, solves the row shard
[g*S, (g+1)*S) of the side being updated, then every replica gets the solved shard -- the exchange Flink does with a
join + coGroup network shuffle between half-steps (Appendix A.1 `updateFactors` message build).  Per-row arithmetic is
identical to the 1-GPU run, so results are bitwise equal.  How the shard travels:
  * `ContextALSBackend.enable_p2p` (rank 64): the kernel itself stores each solved plane row into every rank's replica over
    NVLink (CUDA IPC mappings) and the half-step kernels order themselves with epoch flags kept in the replicas -- no
    collective and no host barrier between half-steps;
  * otherwise an NCCL all-gather of the solved shard (optionally in `max_chunks` chunks on a second stream; default 1: the
    persistent solver kernel owns every SM, so the NCCL kernel cannot really run beside it).

SVM: rank g sweeps the CoCoA blocks k with k % G == g; the per-rank sum of block deltas (dim floats) is summed over the ranks
and every rank applies w += (stepsize / blocks) * delta -- FlinkML's broadcast of w and union+reduce of the K delta vectors
(Appendix A.2) -- either by ONE fused kernel over IPC-mapped exchange buffers (`ContextSVMBackend.enable_p2p`) or by an NCCL
all-reduce followed by the apply kernel.

The compute backend is any object with the small interface used below; the product backend is
`_native.Context` (CUDA kernels through the C-ABI).  torch is plumbing only: device views + collectives.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def device_view(ptr: int, numel: int, dtype=torch.float32, device="cuda:0") -> torch.Tensor:
    """Zero-copy torch view of `numel` elements at raw device address `ptr` (memory owned by the fms_ctx)."""
    typestr = {torch.float32: "<f4", torch.float64: "<f8", torch.int32: "<i4", torch.int64: "<i8", torch.float16: "<f2"}[dtype]

    class _Iface:
        pass

    w = _Iface()
    w.__cuda_array_interface__ = {"shape": (int(numel),), "typestr": typestr, "data": (int(ptr), False), "version": 2}
    return torch.as_tensor(w, device=device)


def shard_range(n: int, world: int, rank: int):
    """Rows [begin, end) of rank `rank`: S = ceil(n / world) rows per shard, last shards may be short or empty."""
    s = (n + world - 1) // world
    b = min(n, s * rank)
    return b, min(n, b + s), s


def sharded_upload(host_arrays, device, group=None):
    """Each rank copies 1/G of every host array (pinned torch tensors of equal length) to its GPU and the slices are
    all-gathered over NVLink: every rank ends up with the full arrays in HBM having pulled only 1/G over PCIe."""
    world, me = dist.get_world_size(group), dist.get_rank(group)
    out = []
    for h in host_arrays:
        n = h.numel()
        s = (n + world - 1) // world
        full = torch.empty(s * world, dtype=h.dtype, device=device)
        b, e = min(n, s * me), min(n, s * (me + 1))
        mine = full[s * me: s * (me + 1)]
        mine[: e - b].copy_(h[b:e], non_blocking=True)
        dist.all_gather_into_tensor(full, mine, group=group)
        out.append(full[:n])
    return out


class ContextALSBackend:
    """Adapts a `_native.Context` (ratings already loaded with this rank's shard) to DistributedALS."""

    def __init__(self, ctx, device):
        self.ctx = ctx
        self.device = device
        self.stream = torch.cuda.ExternalStream(ctx.stream, device=device)
        self.p2p = False

    def half_step(self, which, lam):
        self.ctx.als_half_step(which, lam)

    def half_step_chunk(self, which, lam, chunk, n_chunks):
        self.ctx.als_half_step_chunk(which, lam, chunk, n_chunks)

    def factor_tensor(self, which, rank):
        ptr, rows_padded, b, e = self.ctx.als_factors_device(which)
        return device_view(ptr, rows_padded * rank, torch.float32, self.device).view(rows_padded, rank)

    def updated(self, which):
        self.ctx.als_factors_updated(which)

    def _planes_supported(self):
        """False when the tensor-core path (and so the gather planes) does not cover this ctx's rank: the one
        condition under which falling back to the NCCL all-gather is legitimate.  Every other error propagates."""
        from ._native import ERR_UNSUPPORTED, FlinkMSError
        try:
            self.ctx.als_planes_device(0)
        except FlinkMSError as e:
            if e.status != ERR_UNSUPPORTED:
                raise
            return False
        return True

    def enable_p2p(self, group=None):
        """Map every rank's gather planes into this process (CUDA IPC) so that the half-step kernel writes its plane rows
        into all replicas itself.  Returns False only when the tensor-core path (and so the planes) is not in use on some
        rank; a failing IPC exchange raises.  The mappings belong to the ctx's plane buffers: when they are already in
        place for this world size (same geometry re-loaded) nothing is exchanged again."""
        world = dist.get_world_size(group)
        ok = torch.tensor([1.0 if self._planes_supported() else 0.0], device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)           # all ranks or none
        if ok.item() == 0:
            self.p2p = False
            return False
        mapped = torch.tensor([1.0 if all(self.ctx.als_planes_peer_count(w) == world for w in (0, 1)) else 0.0], device=self.device)
        dist.all_reduce(mapped, op=dist.ReduceOp.MIN, group=group)
        if mapped.item() == 0:
            mine = torch.frombuffer(bytearray(self.ctx.als_planes_ipc_handle(0) + self.ctx.als_planes_ipc_handle(1)), dtype=torch.uint8).to(self.device)
            table = torch.empty(world * 128, dtype=torch.uint8, device=self.device)
            dist.all_gather_into_tensor(table, mine, group=group)        # one 128-byte all-gather, no pickling
            table = table.cpu().numpy().reshape(world, 2, 64)
            for which in (0, 1):
                self.ctx.als_planes_open_peers(which, table[:, which, :].tobytes(), world)
            # fresh flag words on every rank before anybody's first half-step can signal into them
            dist.all_reduce(torch.zeros(1, device=self.device), group=group)
            torch.cuda.synchronize(self.device)
        # the half-step kernels order themselves across the ranks (epoch flags in the replicas): no collective between them
        self.ctx.als_planes_flag_sync(True)
        self.p2p = True
        return True

    def plane_tensors(self, which, rank):
        """fp16 views of the gather planes as a tuple of tensors to all-gather, or None when the tensor-core path does
        not cover `rank`.  The planes are interleaved (one row = rank hi halves then rank lo halves), i.e. ONE tensor."""
        if not self._planes_supported():
            return None
        hi, _lo, rows = self.ctx.als_planes_device(which)
        return (device_view(hi, rows * 2 * rank, torch.float16, self.device).view(rows, 2 * rank),)

    def planes_updated(self, which):
        self.ctx.als_planes_updated(which)


class DistributedALS:
    """iterate(T){ users = f(items); items = f(users) }; users = f(items), sharded by rows."""

    def __init__(self, backend, rank_dim: int, group=None, max_chunks: int = 1, min_chunk_rows: int = 1024):
        self.b = backend
        self.rank_dim = rank_dim
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._stale = set()
        self.max_chunks = max_chunks
        self.min_chunk_rows = min_chunk_rows
        self._comm = None
        self._scratch = None
        self._tok = None

    def _gather(self, full):
        s = full.shape[0] // self.world
        me = dist.get_rank(self.group)
        stream = getattr(self.b, "stream", None)
        if stream is not None:
            with torch.cuda.stream(stream):                              # order the collective after the solve kernel
                dist.all_gather_into_tensor(full, full[me * s:(me + 1) * s], group=self.group)
        else:
            dist.all_gather_into_tensor(full, full[me * s:(me + 1) * s].clone(), group=self.group)

    def exchange(self, which):
        """All-gather what the next half-step reads.  With the tensor-core kernel that is the pair of fp16 gather
        planes (same bytes as the fp32 rows, no re-split afterwards); the fp32 matrix is then gathered once in
        finalize().  Otherwise the fp32 matrix itself."""
        if self.world == 1:
            return
        planes = self.b.plane_tensors(which, self.rank_dim) if hasattr(self.b, "plane_tensors") else None
        if planes is not None:
            for t in planes:
                self._gather(t)
            self.b.planes_updated(which)
            self._stale.add(which)
        else:
            self._gather(self.b.factor_tensor(which, self.rank_dim))
            self.b.updated(which)

    def finalize(self):
        """Bring every replica of the fp32 factor matrices up to date (needed before factors are read out)."""
        for which in sorted(self._stale):
            self._gather(self.b.factor_tensor(which, self.rank_dim))
        self._stale.clear()

    def refresh_replicas(self):
        """All-gather both fp32 factor matrices unconditionally (after a fit that skipped finalize, before serving predict)."""
        if self.world == 1:
            return
        self._stale.update((0, 1))
        self.finalize()

    def _token(self):
        if self._tok is None:
            self._tok = torch.zeros(1, device=self.b.device)
        return self._tok

    def _overlap_plan(self, which):
        """(planes, K) when the half-step can be pipelined: K equal chunks of the shard, chunk k's all-gather running on
        a second stream while chunk k+1 is being solved."""
        if self.world == 1 or self.max_chunks < 2 or not hasattr(self.b, "half_step_chunk") or getattr(self.b, "stream", None) is None:
            return None, 1
        planes = self.b.plane_tensors(which, self.rank_dim) if hasattr(self.b, "plane_tensors") else None
        if planes is None:
            return None, 1
        s = planes[0].shape[0] // self.world
        k = next((k for k in range(self.max_chunks, 1, -1) if s % k == 0 and s // k >= self.min_chunk_rows), 1)
        return planes, k

    def half_step(self, which, lam):
        if self.world > 1 and getattr(self.b, "p2p", False):
            # the kernel stored every solved plane row into all replicas over NVLink and published its epoch there; the next
            # half-step waits for every rank's epoch inside the kernel (fms_als_planes_flag_sync): nothing to do on the host
            self.b.half_step(which, lam)
            self.b.planes_updated(which)
            self._stale.add(which)
            return
        planes, K = self._overlap_plan(which)
        if K == 1:
            self.b.half_step(which, lam)
            self.exchange(which)
            return
        full = planes[0]                                                 # [world * S, 2 * rank] fp16
        S = full.shape[0] // self.world
        c = S // K
        me = dist.get_rank(self.group)
        if self._comm is None:
            self._comm = torch.cuda.Stream(device=full.device)
        if self._scratch is None or self._scratch.shape != (self.world * c, full.shape[1]):
            self._scratch = torch.empty((self.world * c, full.shape[1]), dtype=full.dtype, device=full.device)
        grid = full.view(self.world, S, full.shape[1])
        for k in range(K):
            self.b.half_step_chunk(which, lam, k, K)                     # solver kernel on the library's stream
            ev = torch.cuda.Event()
            ev.record(self.b.stream)
            with torch.cuda.stream(self._comm):
                self._comm.wait_event(ev)
                dist.all_gather_into_tensor(self._scratch, full[me * S + k * c: me * S + (k + 1) * c], group=self.group)
                grid[:, k * c:(k + 1) * c].copy_(self._scratch.view(self.world, c, full.shape[1]))
        done = torch.cuda.Event()
        done.record(self._comm)
        self.b.stream.wait_event(done)                                   # the next half-step reads the gathered planes
        self.b.planes_updated(which)
        self._stale.add(which)

    def iteration(self, lam):
        self.half_step(0, lam)
        self.half_step(1, lam)

    def fit(self, iterations, lam, finalize=True):
        """finalize=False leaves the fp32 matrices sharded (every rank's own shard is current): enough when each rank
        emits its own rows, like the parallel sinks of the reference job."""
        for _ in range(iterations):
            self.iteration(lam)
        self.half_step(0, lam)
        if finalize:
            self.finalize()


class ContextSVMBackend:
    def __init__(self, ctx, device):
        self.ctx = ctx
        self.device = device
        self.stream = torch.cuda.ExternalStream(ctx.stream, device=device)

    def sweep(self, local_iterations):
        self.ctx.svm_sweep(local_iterations)

    def delta_tensor(self):
        ptr, dim = self.ctx.svm_delta_device()
        return device_view(ptr, dim, torch.float32, self.device)

    def apply(self):
        self.ctx.svm_apply_delta()

    p2p = False

    def enable_p2p(self, group=None):
        """Map every rank's delta-w exchange buffer into this process (CUDA IPC): a round is then the sweep plus ONE kernel
        that pushes the rank's delta into all buffers over NVLink, waits for everybody's epoch flag and applies the rank-ordered
        sum -- no collective call between rounds.  A failing IPC exchange raises."""
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        mine = torch.frombuffer(bytearray(self.ctx.svm_exchange_ipc_handle(world)), dtype=torch.uint8).to(self.device)
        table = torch.empty(world * 64, dtype=torch.uint8, device=self.device)
        dist.all_gather_into_tensor(table, mine, group=group)
        self.ctx.svm_exchange_open_peers(table.cpu().numpy().tobytes(), world, rank)
        dist.all_reduce(torch.zeros(1, device=self.device), group=group)     # fresh flag words everywhere before the first round
        torch.cuda.synchronize(self.device)
        self.p2p = True
        return True

    def round_exchange(self, local_iterations):
        self.ctx.svm_round_exchange(local_iterations)


class DistributedSVM:
    """CoCoA outer round: local block sweeps -> all-reduce(sum) of the delta -> w += scaling * delta."""

    def __init__(self, backend, group=None):
        self.b = backend
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1

    def round(self, local_iterations):
        if self.world > 1 and getattr(self.b, "p2p", False):
            self.b.round_exchange(local_iterations)          # sweep + fused push / flag wait / rank-ordered sum / apply
            return
        self.b.sweep(local_iterations)
        if self.world > 1:
            delta = self.b.delta_tensor()
            stream = getattr(self.b, "stream", None)
            if stream is not None:
                with torch.cuda.stream(stream):
                    dist.all_reduce(delta, op=dist.ReduceOp.SUM, group=self.group)
            else:
                dist.all_reduce(delta, op=dist.ReduceOp.SUM, group=self.group)
        self.b.apply()

    def fit(self, iterations, local_iterations):
        for _ in range(iterations):
            self.round(local_iterations)
