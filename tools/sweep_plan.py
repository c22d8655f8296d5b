"""Sweeps the block-gradient GEMM's tile shape (MH) and split-K factor through the SMT_GEMM_FORCE_* environment
knobs and prints the measured time of every combination next to the planner's own choice — the data the
planner's cost-model constants are fitted on (profiles/r01_plan_sweep.md)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sparse_matrix_tuning_b200 import ops

flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=7, warmup=2):
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


CASES = [(256, 8192, 1), (256, 8192, 4), (256, 8192, 9), (256, 8192, 18), (256, 8192, 31), (256, 8192, 60),
         (256, 8192, 100), (256, 8192, 148), (256, 8192, 296), (256, 16384, 13), (256, 16384, 45), (256, 2048, 13),
         (256, 2048, 3), (128, 16384, 51), (128, 8192, 10), (64, 16384, 204), (64, 8192, 40)]
if len(sys.argv) > 1:
    CASES = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for b, T, n in CASES:
    fin, fout = 4096, 4096 if n <= (4096 // b) ** 2 else 4096 * 4
    x = torch.randn(T, fin, device="cuda").bfloat16()
    dy = torch.randn(T, fout, device="cuda").bfloat16()
    perm = torch.randperm((fout // b) * (fin // b))[:n]
    rc = ops.make_block_rc([(int(p) // (fin // b), int(p) % (fin // b)) for p in perm], "cuda")
    out = torch.empty(n * b, b, device="cuda", dtype=torch.bfloat16)
    os.environ.pop("SMT_GEMM_FORCE_MH", None); os.environ.pop("SMT_GEMM_FORCE_SPLITS", None)
    s0, c0 = ops.block_grad_gemm_plan(n, b, T, torch.bfloat16)
    t0 = timeit(lambda: ops.block_grad_gemm(x, dy, rc, b, out=out))
    fl = 2.0 * b * b * T * n
    print(f"## b={b} T={T} n={n}: planner splits={s0} ctas={c0}: {t0:.1f} us {fl / t0 / 1e6:.0f} TF/s", flush=True)
    kt = (T + 63) // 64
    rows = []
    for mh in ((1, 2) if b == 256 else (1,)):
        for s in (1, 2, 3, 4, 6, 8, 11, 16, 24, 32):
            if s > max(1, kt // 4):
                continue
            os.environ["SMT_GEMM_FORCE_MH"] = str(mh)
            os.environ["SMT_GEMM_FORCE_SPLITS"] = str(s)
            se, ce = ops.block_grad_gemm_plan(n, b, T, torch.bfloat16)
            if se != s or ce > 148 * 6:
                continue
            t = timeit(lambda: ops.block_grad_gemm(x, dy, rc, b, out=out))
            rows.append((t, mh, s, ce))
    rows.sort()
    print("   best: " + "  ".join(f"[mh={mh} s={s} ctas={ce}: {t:.1f}us]" for t, mh, s, ce in rows[:4]))
    print("   all : " + " ".join(f"{mh}/{s}:{t:.0f}" for t, mh, s, ce in sorted(rows, key=lambda r: (r[1], r[2]))), flush=True)
    del x, dy
