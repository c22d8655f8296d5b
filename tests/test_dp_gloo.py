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


class _FakeSink:
    def __init__(self, view):
        self.view, self.touched = view, True


class _FakeArena:
    """What OverlappedGradExchange reads from an SMTAdam arena: the flat buffer, per-parameter views and sinks."""

    def __init__(self, sizes, fill):
        self.flat_grad = torch.full((sum(sizes),), float(fill))
        self.grad_views, self.sinks, off = [], [], 0
        for n in sizes:
            v = self.flat_grad[off:off + n]
            self.grad_views.append(v)
            self.sinks.append(_FakeSink(v))
            off += n
        self.all_sinks = True
        self.sq_override, self.sq_dirty = None, False

    def reconcile_grads(self):
        pass


class _FakeArenaOptimizer:
    def __init__(self, arena):
        self._arenas, self.grad_scale = [arena], 1.0


class _LooseOptimizer:
    """SMTAdam(flatten=False): no flat buffers, gradients live on the parameters."""

    def __init__(self, params):
        self.param_groups, self.grad_scale = [{"params": params}], 1.0
        self.modified = False

    def flat_grads(self):
        return []

    def mark_grads_modified(self):
        self.modified = True


class _FakeActAccumulator:
    def __init__(self, acc):
        self.acc = acc


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
        # 4. unflattened optimizer: the per-parameter gradients are reduced (never a scale without a reduction)
        ps = [torch.nn.Parameter(torch.zeros(4)), torch.nn.Parameter(torch.zeros(2))]
        ps[0].grad = torch.full((4,), float(rank + 1))
        ps[1].grad = torch.full((2,), 10.0 * (rank + 1))
        loose = _LooseOptimizer(ps)
        dp.allreduce_compact_grads(loose)
        assert loose.grad_scale == 0.5 and loose.modified
        assert torch.equal(ps[0].grad, torch.full((4,), 3.0)) and torch.equal(ps[1].grad, torch.full((2,), 30.0))
        # 5. activation warm-up: only the batch-reduced [S, C] accumulators cross ranks, SUM like fine_tune.py:655
        act = _FakeActAccumulator({("q_proj", 0): torch.full((3, 4), float(rank + 1)),
                                   ("down_proj", 1): torch.full((3, 8), 2.0 * (rank + 1))})
        dp.allreduce_activation_scores(act)
        assert torch.equal(act.acc[("q_proj", 0)], torch.full((3, 4), 3.0))
        assert torch.equal(act.acc[("down_proj", 1)], torch.full((3, 8), 6.0))
        # 6. overlapped exchange: chunks reduced as they are flushed, every view exactly once, stragglers at finish()
        arena = _FakeArena([8, 16, 8, 24], fill=rank + 1)
        ex = dp.OverlappedGradExchange(_FakeArenaOptimizer(arena))
        try:
            ex._on_flush([arena.sinks[3], arena.sinks[2]])            # backward order: last parameters first
            assert ex.reduced_ranges == [(24, 32)]                    # two adjacent views merged into one collective
            assert torch.equal(arena.flat_grad[24:], torch.full((32,), 3.0)) and arena.flat_grad[0] == rank + 1
            ex._on_flush([arena.sinks[0]])
            ex.finish()                                               # sinks[1] never went through a grouped flush
            assert torch.equal(arena.flat_grad, torch.full((56,), 3.0))
            assert arena.sq_dirty and arena.sq_override is None
            # gradient accumulation: nothing is exchanged until the last micro-batch
            arena.flat_grad.fill_(float(rank + 1))
            ex.active = False
            ex._on_flush(arena.sinks)
            assert arena.flat_grad[0] == rank + 1
            ex.active = True
            ex._on_flush(arena.sinks)
            ex.finish()
            assert torch.equal(arena.flat_grad, torch.full((56,), 3.0))
        finally:
            ex.close()
        # 7. bit-level replica check
        class _A:
            pass
        a = _A()
        a.flat_param = torch.arange(8, dtype=torch.float32).bfloat16()
        a.master = torch.arange(8, dtype=torch.float32)
        o = _A()
        o._arenas = [a]
        assert dp.replicas_identical(o)
        if rank == 1:
            a.master[3] = torch.nextafter(a.master[3], torch.tensor(10.0))   # one ulp off on one rank
        assert not dp.replicas_identical(o)
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
