"""A/B of the split-K reduction of the block-gradient GEMM: thread-block cluster + DSMEM (default) vs the global
workspace paths (SMT_GEMM_CLUSTER_REDUCE=0), on per-module launches of BASELINE config 2 (b = 256 and smaller blocks).
CUDA events, 256 MB memset between iterations, median of 9.  Measurement tooling for an experiment that was NOT kept: the
`SMT_GEMM_CLUSTER_REDUCE` switch only exists with `profiles/r02_cluster_dsmem_reduce_experiment.patch` applied (results:
`profiles/r02_cluster_dsmem_reduce.md`)."""
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
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2] * 1e3


print("| weight | b | n | T | global-workspace reduce us (splits) | cluster / DSMEM reduce us (splits) | speed-up |")
print("|---|---:|---:|---:|---:|---:|---:|")
g = torch.Generator().manual_seed(7)
for fout, fin, label in ((4096, 4096, "q"), (14336, 4096, "gate/up")):
    for T in (2048, 8192, 16384):
        x = torch.randn(T, fin, device="cuda").bfloat16()
        dy = torch.randn(T, fout, device="cuda").bfloat16()
        for b, ns in ((256, (1, 2, 5, 12, 17, 44)), (128, (5, 20, 51)), (64, (20, 81))):
            for n in ns:
                total = (fout // b) * (fin // b)
                if n > total:
                    continue
                perm = torch.randperm(total, generator=g)[:n]
                idx = [(int(p) // (fin // b), int(p) % (fin // b)) for p in perm]
                rc = ops.make_block_rc(idx, "cuda")
                out = torch.empty(n * b, b, device="cuda", dtype=torch.bfloat16)
                res = []
                for mode in ("0", "1"):
                    os.environ["SMT_GEMM_CLUSTER_REDUCE"] = mode
                    t = timeit(lambda: ops.block_grad_gemm(x, dy, rc, b, out=out))
                    res.append((t, ops.block_grad_gemm_plan(n, b, T, torch.bfloat16)[0]))
                print(f"| {label} | {b} | {n} | {T} | {res[0][0]:.1f} ({res[0][1]}) | {res[1][0]:.1f} ({res[1][1]}) | "
                      f"{res[0][0] / res[1][0]:.2f}x |", flush=True)
        del x, dy
