"""GPU tests at BASELINE.json's full sizes, through size-independent properties (the CPU oracle cannot finish
these sizes in seconds): linearity and token-additivity of the block-gradient contraction, split-K invariance,
gather->scatter round trips, optimizer idempotence under re-partitioning, top-k sortedness."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(built_library):
    from sparse_matrix_tuning_b200 import ops as _ops
    return _ops


def _blocks(fout, fin, b, n, seed):
    g = torch.Generator().manual_seed(seed)
    perm = torch.randperm((fout // b) * (fin // b), generator=g)[:n]
    return [(int(p) // (fin // b), int(p) % (fin // b)) for p in perm]


@pytest.mark.parametrize("fin,fout,block,n", [(4096, 4096, 256, 13), (4096, 14336, 256, 45), (14336, 4096, 128, 20),
                                              (4096, 4096, 64, 100)])
def test_gemm_full_size_properties(ops, fin, fout, block, n):
    """Config 2 sizes (4096x4096, 14336x4096, 4096x14336 weights, seq 2048 x batch 4 = 8192 tokens, bf16)."""
    T = 8192
    torch.manual_seed(fin + block)
    x = torch.randn(T, fin, device="cuda").bfloat16()
    dy1 = torch.randn(T, fout, device="cuda").bfloat16()
    rc = ops.make_block_rc(_blocks(fout, fin, block, n, seed=n), "cuda")
    G = ops.block_grad_gemm(x, dy1, rc, block, out_dtype=torch.float32)
    scale = G.abs().max().item()
    # (1) spot-check three blocks against torch fp32 matmul (fp32 reference kept for the floating-point kernel)
    host_rc = rc.cpu().tolist()
    for i in (0, n // 2, n - 1):
        r, c = host_rc[i]
        ref = dy1[:, r * block:(r + 1) * block].float().t() @ x[:, c * block:(c + 1) * block].float()
        assert (G[i * block:(i + 1) * block] - ref).abs().max().item() <= 2e-5 * scale
    # (2) additivity over tokens: G(T) == G(first half) + G(second half)  (exercises different split-K plans)
    h = T // 2 + 64
    Ga = ops.block_grad_gemm(x[:h], dy1[:h], rc, block, out_dtype=torch.float32)
    Gb = ops.block_grad_gemm(x[h:], dy1[h:], rc, block, out_dtype=torch.float32)
    assert (G - (Ga + Gb)).abs().max().item() <= 2e-5 * scale
    # (3) exact linearity in dy for power-of-two scaling, and sign symmetry
    G2 = ops.block_grad_gemm(x, dy1 * 2, rc, block, out_dtype=torch.float32)
    assert torch.equal(G2, G * 2)
    Gn = ops.block_grad_gemm(x, -dy1, rc, block, out_dtype=torch.float32)
    assert torch.equal(Gn, -G)
    # (4) permuting the block list permutes the output rows exactly
    perm = torch.randperm(n).tolist()
    rc_p = rc[perm].contiguous()
    Gp = ops.block_grad_gemm(x, dy1, rc_p, block, out_dtype=torch.float32)
    assert torch.equal(Gp.view(n, block, block), G.view(n, block, block)[perm])
    # (5) bf16 output is the RNE rounding of the fp32 output of the same plan
    Gbf = ops.block_grad_gemm(x, dy1, rc, block, out_dtype=torch.bfloat16)
    assert torch.equal(Gbf, G.bfloat16())


def test_llama8b_compact_state_round_trip_and_adam_repartition(ops):
    """LLaMA-3-8B at 0.71 %: 869 blocks of 256x256 spread over q/k/v-shaped weights. gather(scatter(c)) == c, and an
    Adam step over the whole flat state equals the same step run in two halves (elementwise kernel, any partition)."""
    b, n = 256, 869
    torch.manual_seed(0)
    Wq = torch.randn(4096, 4096, device="cuda").bfloat16()
    Wk = torch.randn(1024, 4096, device="cuda").bfloat16()
    ents = []
    for i in range(n):
        w = Wq if i % 3 else Wk
        j = i * 7919
        ents.append((w, j % (w.shape[0] // b), (j // 16) % (w.shape[1] // b)))
    ents = list({(id(w), r, c): (w, r, c) for w, r, c in ents}.values())   # no duplicate targets
    n = len(ents)
    tab = ops.make_block_table(ents, "cuda")
    comp = torch.randn(n * b, b, device="cuda").bfloat16()
    ops.block_scatter(tab, n, b, comp)
    back = torch.empty_like(comp)
    ops.block_gather(tab, n, b, back)
    assert torch.equal(back, comp)
    N = n * b * b
    g = (torch.randn(N, device="cuda") * 0.01).bfloat16()
    state = [comp.float().reshape(-1).clone(), torch.zeros(N, device="cuda"), torch.zeros(N, device="cuda")]
    whole = [t.clone() for t in state]
    halves = [t.clone() for t in state]
    sq = ops.grad_sqnorm(g)
    kw = dict(lr=1e-4, beta1=0.9, beta2=0.95, eps=1e-8, weight_decay=0.0, step=1, sqnorm=sq, max_norm=1.0)
    ops.compact_adam(*whole, g, table=tab, n_blocks=n, block=b, w_dtype=torch.bfloat16, **kw)
    h = (n // 2) * b * b
    ops.compact_adam(*[t[:h] for t in halves], g[:h], **kw)
    ops.compact_adam(*[t[h:] for t in halves], g[h:], **kw)
    for a, c in zip(whole, halves):
        assert torch.equal(a, c)
    ops.block_gather(tab, n, b, back)                                # fused write-back landed in the dense weights
    assert torch.equal(back.reshape(-1), whole[0].bfloat16())


def test_full_size_scores_and_topk_sorted(ops):
    """One LLaMA-3-8B q_proj-sized accumulator: block scores of a scaled copy scale exactly (power of two), L1 equals
    abs_mean * b^2, and the global top-k comes back sorted by (score, index) descending."""
    torch.manual_seed(1)
    a = torch.randn(4096, 4096, device="cuda")
    s1 = ops.block_score_reduce(a, 256, "L1")
    s2 = ops.block_score_reduce(a * 4, 256, "L1")
    assert torch.equal(s2, s1 * 4)
    am = ops.block_score_reduce(a, 256, "abs_mean")
    assert torch.equal(am, s1 / 65536)
    ma = ops.block_score_reduce(a, 256, "mean_abs")
    assert (ma <= am).all()                                           # |mean| <= mean|.|
    flat = ma.reshape(-1).contiguous()
    idx, _ = ops.topk_blocks(flat, [0, flat.numel()], [64])
    vals = flat[idx.long()].cpu().tolist()
    pairs = list(zip(vals, idx.cpu().tolist()))
    assert pairs == sorted(pairs, reverse=True) and len(set(idx.cpu().tolist())) == 64
    assert min(vals) >= torch.kthvalue(flat.cpu(), flat.numel() - 63).values.item()


def test_selection_at_llama_scale_matches_oracle_exactly(ops):
    """8 LLaMA-3-8B layers of q/k/v accumulators (201 M fp32 elements, 3 072 blocks), n = 217: the drop-in call (host
    tensors in and device tensors in) returns exactly the oracle's selection — keys, order and (row, col) order."""
    from oracle import smt_oracle as O
    from sparse_matrix_tuning_b200.smt import smt_helper as H
    g = torch.Generator().manual_seed(1234)
    dims = {"q_proj": [4096, 4096], "k_proj": [1024, 4096], "v_proj": [1024, 4096]}
    grads = {(m, l): torch.randn(r, c, generator=g) * (1.0 + 0.1 * l) for l in range(8) for m, (r, c) in dims.items()}
    n = int(0.0071 * 122528 * 8 / 32)
    want = O.select_submatrix(grads, dims, n)
    got = H.select_submatrix_based_on_grads(grads, dims, n)
    assert list(got.items()) == list(want.items())
    got_dev = H.select_submatrix_based_on_grads({k: v.cuda() for k, v in grads.items()}, dims, n)
    assert list(got_dev.items()) == list(want.items())


def test_grouped_launch_at_llama_scale_properties(ops):
    """BASELINE config 3 shape: 869 blocks of 256 x 256 over 27 q/k/v-like modules, T = 8192 tokens, ONE grouped launch
    (the cta_group::2 kernel).  Size-independent properties: run-to-run bit-identical, token additivity, exact sign
    symmetry, equality with per-module launches (split-K plans differ => fp32 tolerance), spot checks vs torch fp32."""
    torch.manual_seed(2)
    T, b = 8192, 256
    g = torch.Generator().manual_seed(5)
    shapes = [(4096, 4096)] * 9 + [(1024, 4096)] * 18                 # (out, in): q-like and k/v-like
    xs = [torch.randn(T, 4096, device="cuda").bfloat16() for _ in range(9)]            # q/k/v of a layer share x
    mods, total = [], 0
    for m, (fo, fi) in enumerate(shapes):
        nb = (fo // b) * (fi // b)
        n = 869 // 27 + (1 if m < 869 % 27 else 0)
        perm = torch.randperm(nb, generator=g)[:n].tolist()
        idx = [(p // (fi // b), p % (fi // b)) for p in perm]
        dy = torch.randn(T, fo, device="cuda").bfloat16()
        mods.append((xs[m % 9], dy, idx, total))
        total += n * b * b
    assert total == 869 * b * b

    def run(t0, t1, sign=1.0):
        out = torch.zeros(total, device="cuda")
        batch = ops.BlockGradBatch()
        for x, dy, idx, off in mods:
            batch.add(x[t0:t1], dy[t0:t1] * sign, idx, out[off:off + len(idx) * b * b].view(-1, b), b)
        assert batch.flush(accumulate=False) == 1
        return out

    G = run(0, T)
    assert torch.equal(G, run(0, T))                                  # deterministic
    assert torch.equal(run(0, T, -1.0), -G)                           # exact sign symmetry
    scale = G.abs().max().item()
    h = T // 2 + 192
    assert (G - (run(0, h) + run(h, T))).abs().max().item() <= 2e-5 * scale      # additivity over tokens
    for m in (0, 8, 9, 26):                                           # per-module launches (other split-K plans)
        x, dy, idx, off = mods[m]
        rc = ops.make_block_rc(idx, "cuda")
        Gm = ops.block_grad_gemm(x, dy, rc, b, out_dtype=torch.float32)
        assert (G[off:off + len(idx) * b * b].view(-1, b) - Gm).abs().max().item() <= 2e-5 * scale
        r, c = idx[len(idx) // 2]
        ref = dy[:, r * b:(r + 1) * b].float().t() @ x[:, c * b:(c + 1) * b].float()
        i = len(idx) // 2
        assert (Gm[i * b:(i + 1) * b] - ref).abs().max().item() <= 2e-5 * scale


def test_grouped_launch_with_attention_and_mlp_blocks_at_llama_scale(ops):
    """BASELINE config 5 shape: 1 052 blocks spread over attention (4096 x 4096, 1024 x 4096) AND MLP weights (gate/up
    14336 x 4096 read the same input, down 4096 x 14336 has its own 14336-wide input), T = 4096 tokens, ONE grouped launch
    with per-block sums of squares and a mix of overwriting and accumulating modules.  Size-independent properties:
    determinism, token additivity, sums of squares = what is stored, spot checks of every module kind vs torch fp32."""
    torch.manual_seed(3)
    T, b = 4096, 256
    g = torch.Generator().manual_seed(9)
    x_attn = [torch.randn(T, 4096, device="cuda").bfloat16() for _ in range(3)]       # q/k/v of a layer share it
    x_mlp = [torch.randn(T, 4096, device="cuda").bfloat16() for _ in range(3)]        # gate/up of a layer share it
    x_down = [torch.randn(T, 14336, device="cuda").bfloat16() for _ in range(3)]
    specs = []
    for layer in range(3):
        specs += [(4096, x_attn[layer], 70), (1024, x_attn[layer], 25), (1024, x_attn[layer], 25),
                  (14336, x_mlp[layer], 80), (14336, x_mlp[layer], 80), (4096, x_down[layer], 70)]
    specs[-1] = (4096, x_down[2], 72)                                                  # 1 052 blocks in total
    mods, total = [], 0
    for m, (fo, x, n) in enumerate(specs):
        fi = x.shape[1]
        perm = torch.randperm((fo // b) * (fi // b), generator=g)[:n].tolist()
        idx = [(p // (fi // b), p % (fi // b)) for p in perm]
        dy = torch.randn(T, fo, device="cuda").bfloat16()
        mods.append((x, dy, idx, total, m % 2 == 0))                                   # every other module overwrites
        total += n * b * b
    n_blocks = total // (b * b)
    assert n_blocks == 1052
    init = torch.randn(total, device="cuda")

    def run(t0, t1, with_sq=True):
        out = init.clone()
        sq = torch.full((2 * n_blocks,), -1.0, device="cuda") if with_sq else None
        batch = ops.BlockGradBatch()
        for x, dy, idx, off, overwrite in mods:
            batch.add(x[t0:t1], dy[t0:t1], idx, out[off:off + len(idx) * b * b].view(-1, b), b, accumulate=not overwrite,
                      sq=sq, sq_slot0=2 * (off // (b * b)) if with_sq else -1)
        assert batch.flush(accumulate=True) == 1
        return out, sq

    G, sq = run(0, T)
    assert ops.LAST_GROUP["cta_group_2"] and ops.LAST_GROUP["emits_sq"]
    G2, sq2 = run(0, T)
    assert torch.equal(G, G2) and torch.equal(sq, sq2)                                 # deterministic
    want_sq = (G.view(n_blocks, 2, -1).double() ** 2).sum(-1).reshape(-1)
    assert torch.allclose(sq.double(), want_sq, rtol=1e-5, atol=0)
    scale = (G - init).abs().max().item()
    h = T // 2 + 64
    Ga, _ = run(0, h, with_sq=False)
    Gb, _ = run(h, T, with_sq=False)
    for x, dy, idx, off, overwrite in mods:                                            # additivity over tokens
        sl = slice(off, off + len(idx) * b * b)
        base = 0.0 if overwrite else init[sl]
        lhs = G[sl] - base
        rhs = (Ga[sl] - base) + (Gb[sl] - base)
        assert (lhs - rhs).abs().max().item() <= 2e-5 * scale
    for m in (0, 1, 3, 5, 17):                                                         # q, k, gate, down, last down
        x, dy, idx, off, overwrite = mods[m]
        i = len(idx) // 2
        r, c = idx[i]
        ref = dy[:, r * b:(r + 1) * b].float().t() @ x[:, c * b:(c + 1) * b].float()
        got = G[off + i * b * b: off + (i + 1) * b * b].view(b, b)
        if not overwrite:
            got = got - init[off + i * b * b: off + (i + 1) * b * b].view(b, b)
        assert (got - ref).abs().max().item() <= 2e-5 * scale
