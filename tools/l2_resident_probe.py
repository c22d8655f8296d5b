"""Is the small-block GEMM bound by DRAM or by the per-CTA pipeline?  Same launch (n blocks of b x b over T tokens)
timed with its operands L2-resident (small matrices, no flush) and L2-cold (256 MB memset before every launch)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sparse_matrix_tuning_b200 import ops

flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timeit(fn, do_flush, iters=9, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if do_flush:
            flush.zero_()
        torch.cuda._sleep(200000)        # ~0.1 ms of GPU spin: the launch below is enqueued before the GPU gets to it
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    g = torch.Generator().manual_seed(7)
    for b, feat, n, T in ((64, 1024, 204, 8192), (64, 1024, 148, 8192), (128, 1024, 51, 8192), (128, 2048, 148, 8192),
                          (256, 2048, 37, 8192), (256, 4096, 148, 4096),
                          # BASELINE config 2 points (operands far larger than L2)
                          (64, 4096, 204, 16384), (128, 4096, 51, 16384), (256, 4096, 12, 16384), (256, 4096, 9, 8192),
                          (64, 4096, 40, 8192), (128, 4096, 10, 8192)):
        x = torch.randn(T, feat, device="cuda").bfloat16()
        dy = torch.randn(T, feat, device="cuda").bfloat16()
        total = (feat // b) ** 2
        perm = torch.randperm(total, generator=g)[:n]
        rc = ops.make_block_rc([(int(p) // (feat // b), int(p) % (feat // b)) for p in perm], "cuda")
        out = torch.empty(n * b, b, device="cuda", dtype=torch.bfloat16)
        fn = lambda: ops.block_grad_gemm(x, dy, rc, b, out=out)
        hot, cold = timeit(fn, False), timeit(fn, True)
        splits, ctas = ops.block_grad_gemm_plan(n, b, T, torch.bfloat16)
        fl = 2.0 * b * b * T * n
        print(f"b={b} n={n} T={T} operands {2 * T * feat * 2 / 1e6:.0f} MB splits={splits} ctas={ctas}: "
              f"L2-hot {hot:.1f} us ({fl / hot / 1e6:.0f} TF/s)   L2-cold {cold:.1f} us ({fl / cold / 1e6:.0f} TF/s)", flush=True)


if __name__ == "__main__":
    main()
