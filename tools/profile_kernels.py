"""Small driver for ncu: a handful of launches of every hot SMT kernel at the sizes DESIGN.md quotes.
Usage (on the GPU box):  python tools/profile_kernels.py [gemm|hbm|all]   (prints CUDA-event timings too)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sparse_matrix_tuning_b200 import ops


ITERS = int(os.environ.get("PROFILE_ITERS", "10"))
WARMUP = int(os.environ.get("PROFILE_WARMUP", "3"))


def timeit(fn, iters=None, warmup=None, flush=None):
    iters = ITERS if iters is None else iters
    warmup = WARMUP if warmup is None else warmup
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush is not None:
            flush.zero_()                      # > L2 sized write between timed iterations
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def gemm_cases(flush):
    cases = [(256, 8192, 31, 4096, 1024, "one LLaMA-3-8B k/v module as selected in bench"),
             (256, 8192, 9, 4096, 4096, "avg module at uniform 0.71%"),
             (256, 16384, 13, 4096, 4096, "config 2: 4096x4096, 5%, B=8 x S=2048"),
             (256, 16384, 45, 4096, 14336, "config 2: 14336x4096, 5%"),
             (256, 16384, 148, 4096, 4096, "one tile per SM"),
             (256, 8192, 869, 4096, 57344, "whole-model 0.71% worth of blocks in one launch"),
             (128, 16384, 51, 4096, 4096, "config 2: b=128, 5%"),
             (64, 16384, 204, 4096, 4096, "config 2: b=64, 5%")]
    for b, T, n, fin, fout, what in cases:
        x = torch.randn(T, fin, device="cuda").bfloat16()
        dy = torch.randn(T, fout, device="cuda").bfloat16()
        perm = torch.randperm((fout // b) * (fin // b))[:n]
        rc = ops.make_block_rc([(int(p) // (fin // b), int(p) % (fin // b)) for p in perm], "cuda")
        out = torch.empty(n * b, b, device="cuda", dtype=torch.bfloat16)
        ms = timeit(lambda: ops.block_grad_gemm(x, dy, rc, b, out=out), flush=flush)
        splits, ctas = ops.block_grad_gemm_plan(n, b, T, torch.bfloat16)
        fl = 2.0 * b * b * T * n
        print(f"gemm b={b} T={T} n={n} splits={splits} ctas={ctas}: {ms * 1e3:8.1f} us {fl / ms / 1e9:8.1f} TFLOP/s  # {what}",
              flush=True)
        del x, dy


def hbm_cases(flush):
    R = C = 16384
    acc = torch.zeros(R, C, device="cuda")
    g = torch.randn(R, C, device="cuda").bfloat16()
    ms = timeit(lambda: ops.score_accumulate(acc, g), flush=flush)
    print(f"score_accumulate {R}x{C} bf16: {ms:.3f} ms {R * C * 10 / ms / 1e6:.0f} GB/s (10 B/elt)")
    ms = timeit(lambda: ops.block_score_reduce(acc, 256, "mean_abs"), flush=flush)
    print(f"block_score_reduce b=256: {ms:.3f} ms {R * C * 4 / ms / 1e6:.0f} GB/s (4 B/elt)")
    sums = torch.zeros(R // 256, C // 256, device="cuda")
    ms = timeit(lambda: ops.block_sum_accumulate(sums, g, 256), flush=flush)
    print(f"block_sum_accumulate b=256 bf16: {ms:.3f} ms {R * C * 2 / ms / 1e6:.0f} GB/s (2 B/elt)")
    del acc, g
    n, b = 869, 256
    N = n * b * b
    W = torch.zeros(4096, 4096 * 4, device="cuda", dtype=torch.bfloat16)
    tab = ops.make_block_table([(W, i // 64, i % 64) for i in range(n)], "cuda")
    master, m, v = (torch.zeros(N, device="cuda") for _ in range(3))
    comp = torch.zeros(N, device="cuda", dtype=torch.bfloat16)
    gr = torch.randn(N, device="cuda").bfloat16()
    sq = ops.grad_sqnorm(gr)
    kw = dict(lr=1e-4, beta1=0.9, beta2=0.95, eps=1e-8, weight_decay=0.0, step=3, sqnorm=sq, max_norm=1.0)
    ms = timeit(lambda: ops.compact_adam(master, m, v, gr, table=tab, n_blocks=n, block=b, w_dtype=torch.bfloat16, **kw),
                flush=flush)
    print(f"compact_adam 869 blocks (+W write-back): {ms:.3f} ms {N * 28 / ms / 1e6:.0f} GB/s (28 B/elt)")
    ms = timeit(lambda: ops.compact_adam(master, m, v, gr, compact_out=comp, table=tab, n_blocks=n, block=b,
                                         w_dtype=torch.bfloat16, **kw), flush=flush)
    print(f"compact_adam 869 blocks (+W +compact): {ms:.3f} ms {N * 30 / ms / 1e6:.0f} GB/s (30 B/elt)")
    ms = timeit(lambda: ops.grad_sqnorm(gr, sq), flush=flush)
    print(f"grad_sqnorm 57M bf16: {ms:.3f} ms {N * 2 / ms / 1e6:.0f} GB/s (2 B/elt)")
    comp2 = torch.empty(n * b, b, device="cuda", dtype=torch.bfloat16)
    ms = timeit(lambda: ops.block_gather(tab, n, b, comp2), flush=flush)
    print(f"block_gather 869 blocks: {ms:.3f} ms {N * 4 / ms / 1e6:.0f} GB/s (2+2 B/elt)")


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    noflush = "--no-flush" in sys.argv
    flush = None if noflush else torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")   # 2x L2
    if which in ("gemm", "all"):
        gemm_cases(flush)
    if which in ("hbm", "all"):
        hbm_cases(flush)
