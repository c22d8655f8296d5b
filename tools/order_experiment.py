"""Does the ORDER of the work items matter for small-block launches?  Times `smt_block_grad_gemm` on the same random
block set in three orders: as selected (random), sorted by (row, col) and sorted by (col, row).  Tiles that share a
dy strip (same row) or an x strip (same col) then sit next to each other in the grid and can share the strip in L2.
Timing experiment only (the output row order follows the index list, so results are permuted, not wrong)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sparse_matrix_tuning_b200 import ops

flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=9, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    g = torch.Generator().manual_seed(1234)
    for fout, fin in ((4096, 4096), (14336, 4096)):
        for T in (8192, 16384):
            x = torch.randn(T, fin, device="cuda").bfloat16()
            dy = torch.randn(T, fout, device="cuda").bfloat16()
            for b in (256, 128, 64):
                total = (fout // b) * (fin // b)
                for sp in (0.01, 0.05):
                    n = max(1, int(sp * total))
                    perm = torch.randperm(total, generator=g)[:n]
                    idx = [(int(p) // (fin // b), int(p) % (fin // b)) for p in perm]
                    out = torch.empty(n * b, b, device="cuda", dtype=torch.bfloat16)
                    res = []
                    for label, order in (("random", idx), ("by row", sorted(idx)),
                                         ("by col", sorted(idx, key=lambda rc: (rc[1], rc[0])))):
                        rc = ops.make_block_rc(order, "cuda")
                        res.append((label, timeit(lambda: ops.block_grad_gemm(x, dy, rc, b, out=out))))
                    fl = 2.0 * b * b * T * n
                    print(f"W {fout}x{fin} T={T} b={b} n={n}: " +
                          "  ".join(f"{l} {t:.1f} us ({fl / t / 1e6:.0f} TF/s)" for l, t in res), flush=True)
            del x, dy


if __name__ == "__main__":
    main()
