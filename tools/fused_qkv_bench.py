"""A/B of the fused dense projections (smt_fused_linear_forward / _dgrad) against the library GEMMs they replace
(three torch.matmul per direction + the two adds autograd inserts between the three input gradients).  CUDA events,
256 MB memset between iterations, median of 9.  Measurement tooling."""
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


print("| shape (T x K -> N) | direction | library us | library TFLOP/s | fused tcgen05 us | fused TFLOP/s | speed-up |")
print("|---|---|---:|---:|---:|---:|---:|")
for T, K, Ns in ((8192, 4096, [4096, 1024, 1024]), (16384, 4096, [4096, 1024, 1024]), (8192, 4096, [4096]),
                 (8192, 4096, [14336, 14336]), (2048, 4096, [4096, 1024, 1024])):
    x = torch.randn(T, K, device="cuda").bfloat16()
    ws = [(torch.randn(n, K, device="cuda") * 0.02).bfloat16() for n in Ns]
    dys = [torch.randn(T, n, device="cuda").bfloat16() for n in Ns]
    flops = 2.0 * T * K * sum(Ns)
    t_lib = timeit(lambda: [torch.matmul(x, w.t()) for w in ws])
    t_fus = timeit(lambda: ops.fused_linear_forward(x, ws))
    print(f"| {T} x {K} -> {'+'.join(map(str, Ns))} | forward | {t_lib:.1f} | {flops / t_lib / 1e6:.0f} | {t_fus:.1f} | "
          f"{flops / t_fus / 1e6:.0f} | {t_lib / t_fus:.2f}x |")

    def lib_dgrad():
        acc = torch.matmul(dys[0], ws[0])
        for dy, w in zip(dys[1:], ws[1:]):
            acc = acc + torch.matmul(dy, w)
        return acc

    t_lib = timeit(lib_dgrad)
    t_fus = timeit(lambda: ops.fused_linear_dgrad(dys, ws))
    print(f"| {T} x {'+'.join(map(str, Ns))} -> {K} | dgrad | {t_lib:.1f} | {flops / t_lib / 1e6:.0f} | {t_fus:.1f} | "
          f"{flops / t_fus / 1e6:.0f} | {t_lib / t_fus:.2f}x |")
