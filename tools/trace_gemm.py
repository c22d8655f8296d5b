"""Per-CTA phase timeline of the block-gradient GEMM (debug facility smt_debug_set_gemm_trace).
Usage: python tools/trace_gemm.py [b,T,n ...]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sparse_matrix_tuning_b200 import _lib, ops

NAMES = ["start", "setup", "first_full", "acc_done", "epi_done", "sib_arrived", "reduced"]
cases = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]] or [(256, 8192, 9), (256, 16384, 13), (256, 8192, 148), (256, 2048, 1)]
lib = _lib.load()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
for b, T, n in cases:
    fin = fout = 4096
    x = torch.randn(T, fin, device="cuda").bfloat16()
    dy = torch.randn(T, fout, device="cuda").bfloat16()
    perm = torch.randperm((fout // b) * (fin // b))[:n]
    rc = ops.make_block_rc([(int(p) // (fin // b), int(p) % (fin // b)) for p in perm], "cuda")
    out = torch.empty(n * b, b, device="cuda", dtype=torch.bfloat16)
    splits, ctas = ops.block_grad_gemm_plan(n, b, T, torch.bfloat16)
    for _ in range(3):
        ops.block_grad_gemm(x, dy, rc, b, out=out)
    trace = torch.zeros(ctas * 8, dtype=torch.int64, device="cuda")
    lib.smt_debug_set_gemm_trace(trace.data_ptr(), ctas)
    flush.zero_()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ops.block_grad_gemm(x, dy, rc, b, out=out); e.record()
    torch.cuda.synchronize()
    lib.smt_debug_set_gemm_trace(None, 0)
    t = trace.view(ctas, 8).cpu().double()
    t0 = t[:, 0].min()
    print(f"## b={b} T={T} n={n} splits={splits} ctas={ctas}: event time {a.elapsed_time(e) * 1e3:.1f} us; "
          f"first CTA start -> last stamp {(t.max() - t0) / 1e3:.1f} us")
    for i, name in enumerate(NAMES):
        col = t[:, i]
        col = col[col > 0]
        if len(col):
            print(f"   {name:12s} min {(col.min() - t0) / 1e3:7.2f}  median {(col.median() - t0) / 1e3:7.2f}  max {(col.max() - t0) / 1e3:7.2f} us")
    del x, dy
