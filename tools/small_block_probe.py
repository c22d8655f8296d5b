"""Three launches for an ncu --set full capture: what bounds the small-block (b = 64 / 128) launches, next to the
one-tile-per-SM b = 256 launch.  Usage: ncu --set full -k regex:block_grad_umma --launch-skip 6 -c 3 python tools/small_block_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from sparse_matrix_tuning_b200 import ops

CASES = [(4096, 4096, 64, 204, 16384), (14336, 4096, 128, 179, 16384), (4096, 4096, 256, 148, 16384)]


def main():
    g = torch.Generator().manual_seed(1234)
    work = []
    for fout, fin, b, n, T in CASES:
        x = torch.randn(T, fin, device="cuda").bfloat16()
        dy = torch.randn(T, fout, device="cuda").bfloat16()
        total = (fout // b) * (fin // b)
        perm = torch.randperm(total, generator=g)[:n]
        rc = ops.make_block_rc([(int(p) // (fin // b), int(p) % (fin // b)) for p in perm], "cuda")
        out = torch.empty(n * b, b, device="cuda", dtype=torch.bfloat16)
        work.append((x, dy, rc, b, out))
    for _ in range(3):                       # 2 warm-up rounds (6 launches), then the captured round
        for x, dy, rc, b, out in work:
            ops.block_grad_gemm(x, dy, rc, b, out=out)
        torch.cuda.synchronize()
    print("done")


if __name__ == "__main__":
    main()
