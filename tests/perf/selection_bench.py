"""LLaMA-3-8B-sized warm-up selection: 96 accumulated q/k/v gradients (805 M fp32 elements, 3.2 GB) -> 869 blocks.
Times (i) the oracle port of the reference (CPU strided reductions + Python heap, smt_helper.py:40-146),
(ii) this repo's drop-in call on the same HOST tensors (H2D inside), (iii) the same call on device-resident
accumulators, (iv) the on-device block-sum path; checks that all four select the same blocks.  Measurement tooling."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

from oracle import smt_oracle as O
from sparse_matrix_tuning_b200 import ops
from sparse_matrix_tuning_b200.smt import smt_helper as H

torch.set_num_threads(os.cpu_count() or 1)
g = torch.Generator().manual_seed(1234)
dims = {"q_proj": [4096, 4096], "k_proj": [1024, 4096], "v_proj": [1024, 4096]}
layers = int(sys.argv[1]) if len(sys.argv) > 1 else 32
grads = {}
for layer in range(layers):
    for mod, (r, c) in dims.items():
        grads[(mod, layer)] = torch.randn(r, c, generator=g) * (1.0 + 0.1 * layer)
n = int(0.0071 * 122528 * layers / 32)
elems = sum(t.numel() for t in grads.values())
print(f"{len(grads)} matrices, {elems / 1e6:.0f} M elements ({elems * 4 / 1e9:.2f} GB fp32), n = {n}, host cores = {os.cpu_count()}")

t0 = time.perf_counter(); want = O.select_submatrix(grads, dims, n); t_ref = time.perf_counter() - t0
torch.cuda.synchronize()
H.select_submatrix_based_on_grads({k: v[:256, :256].clone() for k, v in list(grads.items())[:3]},
                                  {k: [256, 256] for k in dims}, 2)                       # warm the kernels up
torch.cuda.synchronize()
t0 = time.perf_counter(); got_host = H.select_submatrix_based_on_grads(grads, dims, n); torch.cuda.synchronize()
t_host = time.perf_counter() - t0
dev = {k: v.cuda() for k, v in grads.items()}
torch.cuda.synchronize()
t0 = time.perf_counter(); got_dev = H.select_submatrix_based_on_grads(dev, dims, n); torch.cuda.synchronize()
t_dev = time.perf_counter() - t0
sums = {k: torch.zeros(v.shape[0] // 256, v.shape[1] // 256, device="cuda") for k, v in dev.items()}
for k, v in dev.items():
    ops.block_sum_accumulate(sums[k], v, 256)
torch.cuda.synchronize()
t0 = time.perf_counter()
got_bs = H.select_submatrix_from_scores(list(sums), [ops.block_sum_finalize(s, 256) for s in sums.values()], n)
torch.cuda.synchronize()
t_bs = time.perf_counter() - t0
same_host = list(got_host.items()) == list(want.items())
same_dev = list(got_dev.items()) == list(want.items())
overlap = sum(len(set(got_bs[k]) & set(want[k])) for k in want) / max(n, 1)
print(f"| path | seconds | identical selection |\n|---|---:|---|")
print(f"| oracle port of the reference on the host CPUs | {t_ref:.3f} | (reference) |")
print(f"| drop-in call, HOST tensors in (3.2 GB H2D inside) | {t_host:.3f} | {same_host} |")
print(f"| drop-in call, device-resident accumulators | {t_dev:.4f} | {same_dev} |")
print(f"| block-sum accumulators -> finalize -> top-k | {t_bs:.4f} | overlap {overlap:.4f} |")
