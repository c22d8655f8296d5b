"""world_size-2 `gloo` tests (CPU) of the data-parallel host logic in sparse_matrix_tuning_b200/dp.py: the compact
gradient all-reduce with the mean folded into grad_scale, the block-sum all-reduce, and the cross-rank selection check."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _FakeOptimizer:
    def __init__(self, bufs):
        self._bufs, self.grad_scale = bufs, 1.0

    def flat_grads(self):
        return self._bufs


class _FakeAccumulator:
    def __init__(self, tensors):
        self.acc = tensors

    def flat_state(self):
        return torch.cat([t.reshape(-1) for t in self.acc.values()])

    def load_flat_state(self, flat):
        off = 0
        for k, t in self.acc.items():
            t.copy_(flat[off:off + t.numel()].view_as(t))
            off += t.numel()


def _worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sparse_matrix_tuning_b200 import dp
        assert dp.world_size() == world and dp.rank() == rank
        # 1. compact-gradient exchange: SUM all-reduce, mean folded into grad_scale
        g = torch.arange(16, dtype=torch.float32) * (rank + 1)
        opt = _FakeOptimizer([g, torch.ones(8) * (rank + 1)])
        works = dp.allreduce_compact_grads(opt, async_op=True)
        for w in works:
            w.wait()
        assert opt.grad_scale == 1.0 / world
        assert torch.equal(g, torch.arange(16, dtype=torch.float32) * 3)
        assert torch.equal(opt.flat_grads()[1], torch.ones(8) * 3)
        # 2. block-sum exchange -> sums of the MEAN gradient on every rank
        acc = _FakeAccumulator({("q_proj", 0): torch.full((2, 2), float(rank + 1)),
                                ("k_proj", 0): torch.full((1, 2), float(10 * (rank + 1)))})
        dp.allreduce_block_sums(acc)
        assert torch.equal(acc.acc[("q_proj", 0)], torch.full((2, 2), 1.5))
        assert torch.equal(acc.acc[("k_proj", 0)], torch.full((1, 2), 15.0))
        # 3. identical selections pass, diverging ones are caught on every rank
        sel = {("q_proj", 0): [(0, 1), (1, 1)], ("v_proj", 1): [(0, 0)]}
        dp.assert_same_selection(sel)
        bad = dict(sel)
        if rank == 1:
            bad[("q_proj", 0)] = [(1, 1), (0, 1)]     # same set, different ORDER: must be flagged (defines row layout)
        caught = False
        try:
            dp.assert_same_selection(bad)
        except RuntimeError:
            caught = True
        assert caught
        results[rank] = "ok"
    except Exception as e:  # pragma: no cover
        results[rank] = f"{type(e).__name__}: {e}"
    finally:
        dist.destroy_process_group()


def test_dp_exchange_world_size_2():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
        assert dict(results) == {0: "ok", 1: "ok"}


def test_single_process_is_a_noop():
    from sparse_matrix_tuning_b200 import dp
    g = torch.ones(4)
    opt = _FakeOptimizer([g])
    assert dp.allreduce_compact_grads(opt) == [] and opt.grad_scale == 1.0 and torch.equal(g, torch.ones(4))
    dp.assert_same_selection({("q_proj", 0): [(0, 0)]})
