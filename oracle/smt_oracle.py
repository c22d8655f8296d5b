"""CPU oracle for the SMT hot path — TEST INFRASTRUCTURE ONLY.

A plain numpy / torch-CPU restatement of the reference algorithm (yudaohai666/Sparse_Matrix_Tuning),
function by function, each citing the reference lines it follows (paths relative to the reference
root).  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference` legs
may import this module, and only as the checker or the timed CPU baseline — never from the product
package `sparse_matrix_tuning_b200/`, which has no CPU path at all.

Pinning (what makes this oracle trustworthy):
  * selection, block scores, gather/scatter, linearZ forward/backward, budget, freeze/convert/param-group
    logic: pinned against the UNMODIFIED reference modules imported in the build container through
    `oracle/ref_shim.py`; the generated vectors are committed under `tests/golden/` (script:
    `oracle/gen_golden.py`) and `tests/test_oracle_cpu.py` re-checks the oracle against them everywhere,
    and against the live reference whenever `/root/reference` exists.
  * Adam step / clipping (`adamw_fused_step`, `clip_coef`): the arithmetic lives in DeepSpeed
    (`deepspeed==0.16.5`, deepspeed_environment.yml:507; FusedAdam = Apex-derived `multi_tensor_adam`),
    which is NOT vendored in the reference and not installable here.  The reference holds no test or
    golden vector at that boundary => **parity unpinned** for the optimizer arithmetic.  We restate the
    published adam_w_mode update and cross-check it against `torch.optim.AdamW` (fp32) to 1e-6.
"""
from __future__ import annotations

import heapq
import re
from collections import defaultdict

import numpy as np
import torch

BLOCK = 256  # smt.py:22, smt_helper.py:52, fine_tune.py:234 (hard-coded in all three places)

LAYER_RE = re.compile(r"model\.layers\.(\d+)\.")  # fine_tune.py:718, smt.py:90, smt.py:647


# ------------------------------------------------------------------------------------------------------
# block scores and selection
# ------------------------------------------------------------------------------------------------------

def block_scores(grad: torch.Tensor, block: int = BLOCK, strategy: str = "mean_abs",
                 dims=None) -> torch.Tensor:
    """smt_helper.py:55-78 (reshape) + :233-251 (the four reductions). `grad` is an fp32 CPU tensor."""
    rows, cols = (grad.shape if dims is None else dims)
    d1, d2 = int(rows / block), int(cols / block)          # smt_helper.py:57-58
    g = grad.reshape(d1, block, d2, block)                 # smt_helper.py:67
    if strategy == "mean_abs":
        return g.mean(dim=(1, 3)).abs()                    # smt_helper.py:233-235  |mean|
    if strategy == "abs_mean":
        return g.abs().mean(dim=(1, 3))                    # smt_helper.py:238-240
    if strategy == "L1":
        return g.abs().sum(dim=(1, 3))                     # smt_helper.py:243-246
    if strategy == "L2":
        return torch.sqrt(torch.sum(g.abs() ** 2, dim=(1, 3)))  # smt_helper.py:249-251
    raise UnboundLocalError(f"unknown calculate_strategy {strategy!r}")  # smt_helper.py:142 (del of unbound)


def select_from_scores(block_means: dict, n: int, selection_strategy: str = "no_restriction"):
    """smt_helper.py:80-146 given the per-matrix score tensors (dict order = insertion order)."""
    ranked = defaultdict(list)
    if selection_strategy == "norm_dist":                  # smt_helper.py:81-100: n PER MATRIX
        for key, bm in block_means.items():
            idx = torch.argsort(bm.reshape(-1), descending=True)[:n]
            for i in idx:
                ranked[key].append(((i // bm.shape[1]).item(), (i % bm.shape[1]).item()))
        return ranked
    heap = []                                              # smt_helper.py:104: global min-heap
    for key, bm in block_means.items():
        for i in range(bm.shape[0]):
            for j in range(bm.shape[1]):
                item = (bm[i, j].item(), (key, i, j))      # smt_helper.py:114-119
                if len(heap) < n:
                    heapq.heappush(heap, item)
                else:
                    heapq.heappushpop(heap, item)
    heap.sort(reverse=True)                                # smt_helper.py:129-130
    for _score, (key, i, j) in heap:                       # smt_helper.py:138-139
        ranked[key].append((i, j))
    if not heap:
        # smt_helper.py:141-142: `del mean` / `del info` after a loop that never ran (n <= 0 or no blocks)
        raise UnboundLocalError("cannot access local variable 'mean' where it is not associated with a value")
    return ranked


def select_submatrix(grads: dict, targeted_module_dims: dict, n: int,
                     selection_strategy: str = "no_restriction", calculate_strategy: str = "mean_abs",
                     block: int = BLOCK):
    """select_submatrix_based_on_grads, smt_helper.py:40-146, with the block size parameterised
    (the reference hard-codes 256 at :52)."""
    block_means = {}
    for key, grad in grads.items():
        block_means[key] = block_scores(grad, block, calculate_strategy, targeted_module_dims[key[0]])
    return select_from_scores(block_means, n, selection_strategy)


def channel_scores(act: torch.Tensor, strategy: str = "mean_abs") -> torch.Tensor:
    """smt_helper.py:169-183: act is [B, S, C] (accumulated |x|)."""
    a = torch.sum(act.abs(), dim=0)                        # smt_helper.py:170
    if strategy == "mean_abs":
        return torch.mean(a.abs(), dim=0)
    if strategy == "abs_mean":
        return torch.abs(torch.mean(a, dim=0))
    if strategy == "L1":
        return torch.norm(a, p=1, dim=0)
    if strategy == "L2":
        return torch.norm(a, p=2, dim=0)
    raise UnboundLocalError(f"unknown calculate_strategy {strategy!r}")


def select_channels_from_scores(column_means: dict, n: int, selection_strategy: str = "no_restriction"):
    """smt_helper.py:185-230."""
    ranked = defaultdict(list)
    if selection_strategy == "norm_dist":
        for key, cm in column_means.items():
            ranked[key] = torch.argsort(cm, descending=True)[:n].tolist()
        return ranked
    heap = []
    for key, cm in column_means.items():
        for idx in range(cm.shape[0]):
            item = (cm[idx].item(), (key, idx))
            if len(heap) < n:
                heapq.heappush(heap, item)
            else:
                heapq.heappushpop(heap, item)
    heap.sort(reverse=True)
    for _v, (key, idx) in heap:
        ranked[key].append(idx)
    return ranked


def select_channels(activation: dict, n: int, selection_strategy="no_restriction",
                    calculate_strategy="mean_abs"):
    """select_channel_based_on_activation, smt_helper.py:149-230."""
    return select_channels_from_scores({k: channel_scores(a, calculate_strategy) for k, a in activation.items()},
                                       n, selection_strategy)


# ------------------------------------------------------------------------------------------------------
# driver-side arithmetic restated from fine_tune.py
# ------------------------------------------------------------------------------------------------------

TARGET_MODULE_NAMES = ("gate_proj", "up_proj", "down_proj", "q_proj", "k_proj", "v_proj")  # fine_tune.py:217-220


def targeted_module_dims(named_parameters) -> dict:
    """fine_tune.py:221-228: first parameter whose name holds 'weight' and the module kind.
    NOTE the reference iterates a *set* of names (fine_tune.py:217); with disjoint names the result does not
    depend on that order."""
    dims = {}
    for name, p in named_parameters:
        if "weight" in name:
            for t in TARGET_MODULE_NAMES:
                if t in name and t not in dims:
                    dims[t] = [p.shape[0], p.shape[1]]
                    break
    return dims


def block_budget(named_parameters, ratio: float, block: int = BLOCK) -> int:
    """fine_tune.py:231-239: float division, embeddings and lm_head included, int() truncation."""
    total = 0
    for _name, p in named_parameters:
        if isinstance(p, torch.Tensor) and p.ndim == 2:
            total += p.shape[0] / block * p.shape[1] / block
    return int(ratio * total)


def attn_module_name(name: str):
    """fine_tune.py:747 / smt.py:103: q/k/v(/o) dispatch by substring."""
    for m in ("q_proj", "k_proj", "v_proj"):
        if m in name:
            return m
    return None


def warmup_accumulate(acc: dict, named_grads, mlp: bool = False, attention: bool = True) -> dict:
    """fine_tune.py:716-768: sum over steps of the fp32 copies of every q/k/v (and MLP) weight gradient.
    `named_grads` yields (parameter name, gradient tensor)."""
    for name, grad in named_grads:
        m = LAYER_RE.search(name)
        layer = int(m.group(1)) if m else None
        key = None
        if "mlp" in name and mlp:                           # fine_tune.py:723
            mod = "gate_proj" if "gate_proj" in name else "up_proj" if "up_proj" in name else "down_proj"
            key = (mod, layer)
        elif "self_attn" in name and "weight" in name and attention:   # fine_tune.py:744
            mod = attn_module_name(name)
            if mod is not None:
                key = (mod, layer)
        if key is None:
            continue
        g = grad.detach().cpu().to(torch.float32)
        acc[key] = g if key not in acc else acc[key] + g    # fine_tune.py:729-740, 753-765
    return acc


# ------------------------------------------------------------------------------------------------------
# compact <-> dense, linearZ
# ------------------------------------------------------------------------------------------------------

def gather_blocks(weight: torch.Tensor, index_list, block: int = BLOCK) -> torch.Tensor:
    """LinearLayer_MatrixSparsity.__init__, smt.py:312-325."""
    out = torch.empty(len(index_list) * block, block, dtype=weight.dtype)
    for i, (r, c) in enumerate(index_list):
        out[i * block:(i + 1) * block, :] = weight[r * block:(r + 1) * block, c * block:(c + 1) * block]
    return out


def scatter_blocks(weight: torch.Tensor, selected: torch.Tensor, index_list, block: int = BLOCK) -> torch.Tensor:
    """The per-forward write-back, smt.py:332-341 (in place on `weight`)."""
    for i, (r, c) in enumerate(index_list):
        weight[r * block:(r + 1) * block, c * block:(c + 1) * block] = selected[i * block:(i + 1) * block, :]
    return weight


def linearz_forward(x: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """linearZ.forward, smt.py:366: dense y = x @ W^T."""
    return torch.matmul(x, weight.t())


def linearz_backward(x: torch.Tensor, dy: torch.Tensor, weight: torch.Tensor, index_list,
                     block: int = BLOCK):
    """linearZ.backward, smt.py:376-413.  x: [B,S,in], dy: [B,S,out].  Returns (grad_input, grad_weight).
    Arithmetic is done exactly as the reference does it: one batched matmul per block in the INPUT dtype,
    then a sum over the batch dimension in that dtype (bf16 => two roundings per block element)."""
    gw = torch.empty(len(index_list) * block, block, dtype=dy.dtype)
    for i, (r, c) in enumerate(index_list):
        xs = x[:, :, c * block:(c + 1) * block]                                 # smt.py:351-356
        part = torch.matmul(dy.permute(0, 2, 1)[:, r * block:(r + 1) * block, :], xs)   # smt.py:397-403
        gw[i * block:(i + 1) * block, :] = torch.sum(part, dim=0)               # smt.py:404
    gi = torch.matmul(dy, weight)                                               # smt.py:406
    return gi, gw


def linearchannel_backward(x: torch.Tensor, dy: torch.Tensor, weight: torch.Tensor, channel_index_list):
    """linearChannel.forward/backward, smt.py:240-247, 283-286.  x: [B,S,in], dy: [B,S,out].  Returns
    (grad_input, grad_weight [n, out]): one batched matmul in the input dtype, then a sum over the batch."""
    partial = torch.empty(x.shape[0], x.shape[1], len(channel_index_list), dtype=x.dtype)
    for i, index in enumerate(channel_index_list):
        partial[:, :, i] = x[:, :, index]                                       # smt.py:245-247
    gw = torch.sum(torch.matmul(partial.permute(0, 2, 1), dy), dim=0)           # smt.py:283-284
    gi = torch.matmul(dy, weight)                                               # smt.py:286
    return gi, gw


def gather_columns(weight: torch.Tensor, channel_index_list) -> torch.Tensor:
    """Column form of smt.py:198-200 (the reference copies ROWS there — inconsistent with its own gradient, see
    DESIGN.md section 6b): compact[i, :] = W[:, idx[i]]."""
    return torch.stack([weight[:, int(i)] for i in channel_index_list]) if len(channel_index_list) else \
        torch.empty(0, weight.shape[0], dtype=weight.dtype)


def scatter_columns(weight: torch.Tensor, selected: torch.Tensor, channel_index_list) -> torch.Tensor:
    """Column form of smt.py:208-211: W[:, idx[i]] = compact[i, :] (in place)."""
    for i, index in enumerate(channel_index_list):
        weight[:, int(index)] = selected[i, :]
    return weight


def block_grad_truth(x: torch.Tensor, dy: torch.Tensor, index_list, block: int = BLOCK) -> torch.Tensor:
    """fp64 value of the same contraction (the quantity both the reference and the kernel approximate)."""
    x2 = x.reshape(-1, x.shape[-1]).double()
    d2 = dy.reshape(-1, dy.shape[-1]).double()
    out = torch.empty(len(index_list) * block, block, dtype=torch.float64)
    for i, (r, c) in enumerate(index_list):
        out[i * block:(i + 1) * block] = d2[:, r * block:(r + 1) * block].t() @ x2[:, c * block:(c + 1) * block]
    return out


# ------------------------------------------------------------------------------------------------------
# optimizer (DeepSpeed FusedAdam adam_w_mode + global-norm clip) — parity unpinned, see module docstring
# ------------------------------------------------------------------------------------------------------

def bf16_round(a: np.ndarray) -> np.ndarray:
    """fp32 -> bf16 (round-to-nearest-even) -> fp32, on the bit pattern."""
    u = a.astype(np.float32).view(np.uint32).astype(np.uint64)
    rounded = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return rounded.astype(np.uint32).view(np.float32)


def clip_coef(sqnorm: np.float32, grad_scale: np.float32, max_norm: np.float32) -> np.float32:
    """DeepSpeed `clip_tensors_by_global_norm` / torch clip_grad_norm_: coef = max_norm / (norm + 1e-6),
    applied only when < 1 (gradient_clipping: 1.0, helpers/deepspeed_helpers.py:87).
    Returns the multiplier applied to each raw gradient element (including grad_scale)."""
    f = np.float32
    if max_norm <= 0:
        return f(grad_scale)
    norm = f(np.sqrt(f(sqnorm))) * f(grad_scale)
    coef = f(max_norm) / f(norm + f(1e-6))
    return f(grad_scale) * coef if coef < 1 else f(grad_scale)


def partial_sqnorm_total(partials, threads: int = 256) -> np.float32:
    """Total of several partial sums of squares in the FIXED order the Adam kernel uses when it is handed the
    per-block slots of the GEMM epilogue / per-chunk norms of a data-parallel exchange (no reference counterpart: the
    reference's clip, helpers/deepspeed_helpers.py:87, takes one norm over all gradients; the order only fixes the last
    fp32 bits): thread t adds partials[t], partials[t + threads], ... sequentially, each warp of 32 threads is reduced
    by an xor butterfly (16, 8, 4, 2, 1), and the per-warp results by the same butterfly."""
    f = np.float32
    partials = np.asarray(partials, dtype=np.float32).reshape(-1)
    per_thread = np.zeros(threads, dtype=np.float32)
    for t in range(min(threads, len(partials))):
        acc = f(0)
        for val in partials[t::threads]:
            acc = f(acc + val)
        per_thread[t] = acc

    def butterfly(vals):
        vals = vals.astype(np.float32).copy()
        lanes = np.arange(32)
        for o in (16, 8, 4, 2, 1):
            vals = (vals + vals[lanes ^ o]).astype(np.float32)
        return vals[0]

    warps = np.zeros(32, dtype=np.float32)
    for w in range(threads // 32):
        warps[w] = butterfly(per_thread[32 * w:32 * w + 32])
    return f(butterfly(warps))


def adamw_fused_step(p, m, v, g, *, lr, beta1, beta2, eps, weight_decay, step, gscale=1.0):
    """One multi_tensor_adam (adam_w_mode=1, bias_correction=1) update in fp32, every operation rounded
    individually (no FMA contraction):
        g' = g*gscale ; m = b1*m + (1-b1)*g' ; v = b2*v + ((1-b2)*g')*g'
        update = (m/bc1) / (sqrt(v/bc2) + eps) + wd*p ;  p = p - lr*update
    Call site: fine_tune.py:352-363 (`FusedAdam(..., betas=(0.9, 0.95))`), stepped at fine_tune.py:773."""
    f = np.float32
    p, m, v, g = (np.asarray(a, dtype=np.float32) for a in (p, m, v, g))
    b1, b2 = f(beta1), f(beta2)
    bc1, bc2 = f(1.0 - beta1 ** step), f(1.0 - beta2 ** step)
    omb1, omb2 = f(1) - b1, f(1) - b2
    gj = g * f(gscale)
    m = b1 * m + omb1 * gj
    v = b2 * v + (omb2 * gj) * gj
    denom = np.sqrt(v / bc2) + f(eps)
    update = (m / bc1) / denom + f(weight_decay) * p
    p = p - f(lr) * update
    return p.astype(np.float32), m.astype(np.float32), v.astype(np.float32)


# ------------------------------------------------------------------------------------------------------
# CPU modules restating LinearLayer_MatrixSparsity / linearZ — used by bench.py's CPU-baseline legs only
# ------------------------------------------------------------------------------------------------------

class OracleLinearZ(torch.autograd.Function):
    """linearZ, smt.py:347-413, restated on top of linearz_forward / linearz_backward above."""

    @staticmethod
    def forward(ctx, input, selected_weight, matrix_index_list, weight, block):
        ctx.index_list, ctx.block = matrix_index_list, block
        ctx.save_for_backward(input, weight)
        return linearz_forward(input, weight)

    @staticmethod
    def backward(ctx, grad_output):
        x, weight = ctx.saved_tensors
        gi, gw = linearz_backward(x, grad_output, weight, ctx.index_list, ctx.block)
        return gi, gw, None, None, None


class OracleSparseLinear(torch.nn.Module):
    """LinearLayer_MatrixSparsity, smt.py:302-344: gather at construction, scatter at every forward."""

    def __init__(self, weight, index_list, block: int = BLOCK):
        super().__init__()
        self.weight = weight
        self.weight.requires_grad = False
        self.index_list, self.block = index_list, block
        self.selected_weight = torch.nn.Parameter(gather_blocks(weight.data, index_list, block))

    def forward(self, x):
        scatter_blocks(self.weight.data, self.selected_weight.data, self.index_list, self.block)   # smt.py:332-341
        return OracleLinearZ.apply(x, self.selected_weight, self.index_list, self.weight, self.block)
