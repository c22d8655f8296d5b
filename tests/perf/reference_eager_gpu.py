"""Secondary comparator: the reference's OWN algorithm — per-block bmm + sum + slice copy in
backward, per-block slice scatter in forward (restated in oracle/smt_oracle.py) — run as eager PyTorch ON THE B200,
next to this repo's kernels, on the same tensors.  Measurement tooling only (imports oracle/)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from oracle import smt_oracle as O
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


print("| module | B x S | blocks | reference eager on B200: block-grad loop us | this repo: smt_block_grad_gemm us | speed-up | reference scatter loop us | smt_block_scatter us |")
print("|---|---|---:|---:|---:|---:|---:|---:|")
g = torch.Generator().manual_seed(0)
for label, fout, fin, n, B, S in (("q_proj 4096x4096", 4096, 4096, 9, 16, 512), ("k_proj 1024x4096", 1024, 4096, 31, 16, 512),
                                  ("q_proj 4096x4096", 4096, 4096, 13, 8, 2048), ("gate 14336x4096", 14336, 4096, 45, 8, 2048)):
    b = 256
    x = torch.randn(B, S, fin, device="cuda").bfloat16()
    dy = torch.randn(B, S, fout, device="cuda").bfloat16()
    w = (torch.randn(fout, fin, device="cuda") * 0.02).bfloat16()
    perm = torch.randperm((fout // b) * (fin // b), generator=g)[:n]
    idx = [(int(p) // (fin // b), int(p) % (fin // b)) for p in perm]
    rc = ops.make_block_rc(idx, "cuda")
    out = torch.empty(n * b, b, device="cuda", dtype=torch.bfloat16)
    x2, dy2 = x.reshape(-1, fin), dy.reshape(-1, fout)

    def ref_block_loop():                                    # smt.py:382-404 on CUDA tensors
        gw = torch.empty(n * b, b, dtype=dy.dtype, device="cuda")
        for i, (r, c) in enumerate(idx):
            gw[i * b:(i + 1) * b, :] = torch.sum(torch.matmul(dy.permute(0, 2, 1)[:, r * b:(r + 1) * b, :],
                                                              x[:, :, c * b:(c + 1) * b]), dim=0)
        return gw

    t_ref = timeit(ref_block_loop)
    t_ours = timeit(lambda: ops.block_grad_gemm(x2, dy2, rc, b, out=out))
    sel = O.gather_blocks(w.cpu(), idx, b).cuda()
    t_sc_ref = timeit(lambda: O.scatter_blocks(w, sel, idx, b))
    tab = ops.make_block_table([(w, r, c) for r, c in idx], "cuda")
    t_sc = timeit(lambda: ops.block_scatter(tab, n, b, sel))
    err = (ref_block_loop().float() - out.float()).abs().max().item() / out.float().abs().max().item()
    print(f"| {label} | {B} x {S} | {n} | {t_ref:.1f} | {t_ours:.1f} | {t_ref / t_ours:.1f}x | {t_sc_ref:.1f} | {t_sc:.1f} |"
          f"  <!-- max rel diff {err:.2e} -->", flush=True)
