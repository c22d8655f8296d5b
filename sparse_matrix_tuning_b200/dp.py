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
    """SUM-all-reduce the optimizer's flat gradient buffers and fold the 1/world mean into its `grad_scale`."""
    from .smt.smt import flush_block_grads
    flush_block_grads()                                      # the flat buffer must be complete before it is reduced
    ws = world_size(group)
    optimizer.grad_scale = 1.0 / ws
    return allreduce_sum_(optimizer.flat_grads(), group=group, async_op=async_op)


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

    def __init__(self, model, optimizer, group=None):
        self.model, self.optimizer, self.group = model, optimizer, group

    def step(self, loss: torch.Tensor) -> None:
        loss.backward()
        works = allreduce_compact_grads(self.optimizer, group=self.group, async_op=True)
        for w in works:
            w.wait()
        self.optimizer.step()
        self.optimizer.zero_grad()
