"""Grouped launch of row-sharing block pairs: the cta_group::2 kernel (two blocks per SM pair) vs single-CTA tiles.
(An earlier revision also timed a cta_group::1 cluster variant with TMA multicast: profiles/r01_pair_experiment_raw.txt.)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sparse_matrix_tuning_b200 import ops

flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=9, warmup=2):
    """GPU time of the grouped launch(es) only (CUDA events around the C-ABI call inside ops, as bench.py does):
    the host-side descriptor encoding and the pinned copy are outside."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ops.enable_timing("block_grad_gemm")
    ts = []
    for _ in range(iters):
        flush.zero_()
        fn()
        torch.cuda.synchronize()
        ts.append(sum(ms for ms, _tag in ops.collect_timing("block_grad_gemm")))
    ops.enable_timing("block_grad_gemm", False)
    ts.sort()
    return ts[len(ts) // 2] * 1e3


b, T = 256, 8192
g = torch.Generator().manual_seed(0)
for label, n_mod, rows, cols, per_row in (("k/v-like modules: 4x16 blocks, 8 per row", 27, 4, 16, 8),
                                          ("q-like modules: 16x16 blocks, 2 per row", 27, 16, 16, 2),
                                          ("all pairs, 4 per row", 27, 8, 16, 4)):
    xs = [torch.randn(T, cols * b, device="cuda").bfloat16() for _ in range(n_mod)]
    dys = [torch.randn(T, rows * b, device="cuda").bfloat16() for _ in range(n_mod)]
    idxs = []
    for _ in range(n_mod):
        idx = []
        for r in range(rows):
            perm = torch.randperm(cols, generator=g)[:per_row].tolist()
            idx += [(r, c) for c in perm]
        idxs.append(idx)
    n = sum(len(i) for i in idxs)
    out = torch.zeros(n * b * b, device="cuda", dtype=torch.bfloat16)

    def run():
        batch = ops.BlockGradBatch()
        off = 0
        for x, dy, idx in zip(xs, dys, idxs):
            m = len(idx) * b * b
            batch.add(x, dy, idx, out[off:off + m].view(-1, b), b)
            off += m
        batch.flush(accumulate=False)

    fl = 2.0 * b * b * T * n
    os.environ["SMT_GEMM_2SM"] = "1"
    t_2sm = timeit(run)
    os.environ["SMT_GEMM_2SM"] = "0"
    t_single = timeit(run)
    os.environ.pop("SMT_GEMM_2SM", None)
    print(f"{label}: {n} blocks: singles {t_single:.1f} us ({fl / t_single / 1e6:.0f} TF/s)  "
          f"cta_group::2 {t_2sm:.1f} us ({fl / t_2sm / 1e6:.0f} TF/s)",
          flush=True)
    del xs, dys
