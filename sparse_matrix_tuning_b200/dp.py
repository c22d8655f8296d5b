"""Data-parallel exchange for the SMT phase: one process per GPU, `torch.distributed` (NCCL over NVLink/NVSwitch).

SMT shards naturally: every rank holds a full replica (frozen dense weights + compact parameters + optimizer
state) and trains on its own micro-batch.  Only two things ever cross GPUs:

  1. the flat compact-gradient buffer (sum of n_i*b*b elements; 114 MB in bf16 for LLaMA-3-8B at 0.71 %): ONE
     all-reduce per step.  The reference gets this implicitly from the DeepSpeed engine, which reduces the
     gradients of `requires_grad` parameters in 1e6-element buckets during `model.backward()` /
     `model.step()` (fine_tune.py:712,773; helpers/deepspeed_helpers.py:73) — about 57 bucket collectives.
     The 1/world averaging is folded into the Adam kernel's `grad_scale`, so no extra pass touches the buffer.
  2. during warm-up, the per-block signed-sum tensor (12 288 - 98 304 floats) when scores are accumulated in
     block-sum mode.  With identical scores every rank runs the same deterministic top-k, so no index broadcast
     (the reference hands indices around through a file: helpers/deepspeed_helpers.py:177-200).

Everything here is plain `torch.distributed` on whatever tensors it is given, so the host logic is exercised
on CPU with the gloo backend in tests/test_dp_gloo.py.
"""
from __future__ import annotations

import contextlib
import hashlib
from typing import Iterable

import torch
import torch.distributed as dist


def world_size(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def rank(group=None) -> int:
    return dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0


def allreduce_sum_(tensors: Iterable[torch.Tensor], group=None, async_op: bool = False):
    """In-place SUM all-reduce of each tensor; returns the list of work handles when async."""
    if world_size(group) == 1:
        return []
    works = []
    for t in tensors:
        w = dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if async_op:
            works.append(w)
    return works


def allreduce_compact_grads(optimizer, group=None, async_op: bool = False):
    """SUM-all-reduce the optimizer's gradients and fold the 1/world mean into its `grad_scale`: the flat buffers of an
    `SMTAdam(flatten=True)`, or every parameter's `.grad` when the optimizer keeps no flat buffers
    (`flatten=False`) - never a scale without a reduction."""
    from .smt.smt import flush_block_grads
    flush_block_grads()                                      # the flat buffer must be complete before it is reduced
    ws = world_size(group)
    flats = optimizer.flat_grads()
    if ws > 1 and not flats:
        flats = [p.grad for g in optimizer.param_groups for p in g["params"] if p.grad is not None]
        if not flats:
            raise RuntimeError("allreduce_compact_grads: the optimizer holds no gradients to reduce")
    optimizer.grad_scale = 1.0 / ws
    if ws > 1 and hasattr(optimizer, "mark_grads_modified"):
        optimizer.mark_grads_modified()                      # per-block sums of squares are stale after the reduction
    return allreduce_sum_(flats, group=group, async_op=async_op)


class OverlappedGradExchange:
    """All-reduce of the compact gradients OVERLAPPED with the backward pass (the reference gets the
    same effect from DeepSpeed's bucketed reduction during `model.backward()`, fine_tune.py:712,
    helpers/deepspeed_helpers.py:73).

    With `smt.set_grouped_backward(True, chunk_blocks=...)` the block-gradient GEMM is flushed in GPU-filling chunks
    during the backward pass.  After each chunk this object (a flush listener) makes a side stream wait for the launch,
    all-reduces the contiguous range(s) of the flat gradient buffer that chunk produced (NCCL on its own stream) and
    computes that range's sum of squares - all while the main stream keeps running the rest of the backward pass.
    `finish()` (before `optimizer.step()`) joins the side stream and hands the per-chunk sums of squares to the
    optimizer, so neither the exchange nor the clip norm leaves a serialized pass at the end of the step."""

    def __init__(self, optimizer, group=None, sqnorm=None):
        """`sqnorm`: None = compute the per-chunk sums of squares only when there is something to reduce (world > 1; on
        one GPU the GEMM epilogue's own per-block sums are valid); True = always (tests); False = never."""
        from .smt import smt as _smt
        self._smt = _smt
        self.optimizer, self.group = optimizer, group
        self.ws = world_size(group)
        self.sqnorm = (self.ws > 1) if sqnorm is None else bool(sqnorm)
        arenas = [a for a in optimizer._arenas if a is not None]
        if len(arenas) != 1 or not arenas[0].all_sinks:
            raise RuntimeError("OverlappedGradExchange needs an SMTAdam with one flat arena of SMT block parameters")
        self.arena = arenas[0]
        self.flat = self.arena.flat_grad
        self.cuda = self.flat.is_cuda                        # (CPU tensors: gloo tests of the host logic only)
        self.sqnorm = self.sqnorm and self.cuda
        self.side = torch.cuda.Stream(device=self.flat.device) if self.cuda else None
        self._base = self.flat.data_ptr()
        self._esize = self.flat.element_size()
        self._partials = torch.zeros(256, dtype=torch.float32, device=self.flat.device)
        self._n_chunks = 0
        self._reduced = set()                                # data_ptr of every sink view already all-reduced this step
        self.reduced_ranges = []                             # (offset, elements) all-reduced since the last finish()
        self.events = []                                     # (start, end) CUDA events per chunk when profiling
        self.profile = False
        self.active = True          # set False for all but the last micro-batch of a gradient-accumulation step (like DDP.no_sync)
        self.keep_local = None      # debug (bench.py's dp_check): a tensor like `flat` that receives this rank's gradients
                                    # as they were BEFORE each chunk was reduced
        optimizer.grad_scale = 1.0 / self.ws
        self._finished_at = -1                               # optimizer.steps_done at the last finish()
        _smt.add_flush_listener(self._on_flush)
        # `optimizer.step()` joins the exchange by itself if the loop did not call finish() (idempotent per step)
        self._hook = optimizer.register_step_pre_hook(lambda _opt, _a, _k: self.finish()) \
            if hasattr(optimizer, "register_step_pre_hook") else None

    def close(self) -> None:
        self._smt.remove_flush_listener(self._on_flush)
        if self._hook is not None:
            self._hook.remove()
            self._hook = None

    def _ranges(self, sinks):
        spans = sorted(((sk.view.data_ptr() - self._base) // self._esize, sk.view.numel()) for sk in sinks)
        merged = []
        for off, n in spans:
            if merged and merged[-1][0] + merged[-1][1] == off:
                merged[-1][1] += n
            else:
                merged.append([off, n])
        return merged

    def _on_flush(self, sinks) -> None:
        sinks = [sk for sk in sinks if sk.view.data_ptr() >= self._base
                 and sk.view.data_ptr() < self._base + self.flat.numel() * self._esize]
        if not sinks or not self.active:
            return
        if self.cuda:
            main = torch.cuda.current_stream(self.flat.device)
            self.side.wait_stream(main)                      # the chunk's GEMM (just enqueued on `main`) must finish first
            ctx = torch.cuda.stream(self.side)
        else:
            ctx = contextlib.nullcontext()
        with ctx:
            if self.profile and self.cuda:
                e0 = torch.cuda.Event(enable_timing=True)
                e0.record(self.side)
            for off, n in self._ranges(sinks):
                view = self.flat[off:off + n]
                if self.keep_local is not None:
                    self.keep_local[off:off + n].copy_(view)
                if self.ws > 1:
                    w = dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                    w.wait()                                 # CUDA: stream-level, `side` waits for NCCL's stream
                if self.sqnorm:
                    if self._n_chunks >= self._partials.numel():
                        raise RuntimeError("OverlappedGradExchange: more than 256 chunks in one step")
                    from . import ops
                    ops.grad_sqnorm(view, self._partials[self._n_chunks:self._n_chunks + 1])
                    self._n_chunks += 1
                self.reduced_ranges.append((off, n))
            self._reduced.update(sk.view.data_ptr() for sk in sinks)
            if self.profile and self.cuda:
                e1 = torch.cuda.Event(enable_timing=True)
                e1.record(self.side)
                self.events.append((e0, e1))

    def finish(self) -> None:
        """Join the exchange: after this the flat gradient buffer holds the cross-rank SUM on the current stream and the
        optimizer knows the partial sums of squares.  Call between backward and `optimizer.step()`."""
        if not self.active:
            return                                           # exchange suspended (accumulation micro-batch / local-only step)
        serial = getattr(self.optimizer, "steps_done", None)
        if serial is not None and serial == self._finished_at and len(self._smt._pending) == 0 and not self._reduced:
            return                                           # already joined for this gradient state
        self._smt.flush_block_grads()                        # last chunk (normally already flushed by the engine callback)
        if self.cuda:
            torch.cuda.current_stream(self.flat.device).wait_stream(self.side)
        arena = self.arena
        missed = [sk for sk in arena.sinks if sk is not None and sk.touched and sk.view.data_ptr() not in self._reduced]
        if missed:
            # gradients that did not pass through a grouped flush (per-module launches, fp32 path): reduce exactly those
            # views now, the plain way, and let the optimizer recompute the norm over the whole buffer
            for off, n in self._ranges(missed):
                if self.ws > 1:
                    dist.all_reduce(self.flat[off:off + n], op=dist.ReduceOp.SUM, group=self.group)
            arena.sq_override = None
            arena.sq_dirty = True
        elif self.sqnorm and self._n_chunks > 0:
            arena.sq_override = self._partials[:self._n_chunks].clone()
        else:
            arena.sq_dirty = self.ws > 1
        self._n_chunks = 0
        self._reduced = set()
        self.reduced_ranges = []
        self._finished_at = serial if serial is not None else -1


def allreduce_block_sums(accumulator, group=None) -> None:
    """Warm-up, block-sum mode: make every rank hold the block sums of the data-parallel MEAN gradient
    (what `safe_get_full_grad` hands the reference on every rank, fine_tune.py:724)."""
    ws = world_size(group)
    if ws == 1:
        return
    flat = accumulator.flat_state()
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat /= ws
    accumulator.load_flat_state(flat)


def allreduce_activation_scores(accumulator, group=None) -> None:
    """Activation warm-up (K11): make every rank hold the data-parallel SUM of the batch-reduced |x| accumulators.
    The reference all-reduces the FULL [B, S, C] activation of every Linear inside its forward hook
    (fine_tune.py:651-657: `abs()`, `barrier()`, `deepspeed.comm.all_reduce(x)`); the only thing selection consumes
    is sum_b |x| (smt_helper.py:170), which is linear in the per-rank contributions - so one all-reduce of the
    reduced [S, C] accumulators (B times smaller, once per selection instead of once per Linear per step) gives the
    same scores.  `accumulator` is a `warmup.WarmupActivationAccumulator`."""
    ws = world_size(group)
    if ws == 1 or not accumulator.acc:
        return
    keys = list(accumulator.acc.keys())
    flat = torch.cat([accumulator.acc[k].reshape(-1) for k in keys])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for k in keys:
        n = accumulator.acc[k].numel()
        accumulator.acc[k].copy_(flat[off:off + n].view_as(accumulator.acc[k]))
        off += n


def replica_checksum(optimizer) -> torch.Tensor:
    """int64 [2 * groups] bit-level checksums of the flat parameters and fp32 masters: identical on every rank iff the
    replicas hold bit-identical state."""
    sums = []
    for arena in optimizer._arenas:
        if arena is None:
            continue
        sums.append(arena.flat_param.view(torch.int16).to(torch.int64).sum())
        sums.append(arena.master.view(torch.int32).to(torch.int64).sum())
    return torch.stack(sums) if sums else torch.zeros(0, dtype=torch.int64)


def replicas_identical(optimizer, group=None) -> bool:
    """All-gathers `replica_checksum` and compares: True when every rank's parameters and masters are bit-identical."""
    ws = world_size(group)
    if ws == 1:
        return True
    mine = replica_checksum(optimizer)
    gathered = [torch.empty_like(mine) for _ in range(ws)]
    dist.all_gather(gathered, mine, group=group)
    return all(torch.equal(g, gathered[0]) for g in gathered)


def selection_fingerprint(selection: dict) -> str:
    """Order-sensitive digest of a selection dict {(module, layer): [(row, col), ...]}."""
    h = hashlib.sha256()
    for key in sorted(selection.keys(), key=repr):
        h.update(repr((key, list(selection[key]))).encode())
    return h.hexdigest()


def assert_same_selection(selection: dict, group=None) -> None:
    """Cheap cross-rank check that deterministic selection really produced identical index lists."""
    ws = world_size(group)
    if ws == 1:
        return
    digest = bytes.fromhex(selection_fingerprint(selection))
    device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else "cpu"
    mine = torch.tensor(list(digest), dtype=torch.uint8, device=device)
    gathered = [torch.empty_like(mine) for _ in range(ws)]
    dist.all_gather(gathered, mine, group=group)
    for r, other in enumerate(gathered):
        if not torch.equal(other, gathered[0]):
            raise RuntimeError(f"SMT selection differs between rank 0 and rank {r}")


class DataParallelSMT:
    """Minimal native training-step driver for the SMT phase (the part of fine_tune.py:710-773 that touches
    the hot path): backward -> one all-reduce of the compact gradients -> fused Adam step."""

    def __init__(self, model, optimizer, group=None, overlap: bool = False):
        self.model, self.optimizer, self.group = model, optimizer, group
        self.exchange = OverlappedGradExchange(optimizer, group) if overlap else None

    def step(self, loss: torch.Tensor) -> None:
        loss.backward()
        if self.exchange is not None:
            self.exchange.finish()
        else:
            works = allreduce_compact_grads(self.optimizer, group=self.group, async_op=True)
            for w in works:
                w.wait()
        self.optimizer.step()
        self.optimizer.zero_grad()
