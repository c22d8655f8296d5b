"""GPU bring-up script (not a pytest file): exercises every C-ABI kernel against plain torch math and
prints PASS/FAIL per case plus rough timings.  Each group runs in its own subprocess so that a device
trap in one kernel cannot poison the CUDA context of the others.

    python tools/gpu_bringup.py            # all groups
    python tools/gpu_bringup.py gemm       # one group
"""
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

GROUPS = ["score", "topk", "copy", "adam", "gemm_small", "gemm_split", "gemm_f32", "perf"]


def _time_cuda(fn, iters=20, warmup=3):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def report(name, ok, extra=""):
    print(f"[{'PASS' if ok else 'FAIL'}] {name} {extra}", flush=True)
    return ok


def g_score():
    import torch
    from sparse_matrix_tuning_b200 import ops
    torch.manual_seed(0)
    ok = True
    for dt in (torch.bfloat16, torch.float32, torch.float16):
        acc = torch.randn(1000, 1024, device="cuda")
        g = torch.randn(1000, 1024, device="cuda").to(dt)
        ref = acc + g.float()
        ops.score_accumulate(acc, g)
        ok &= report(f"score_accumulate {dt}", torch.equal(acc, ref))
    acc = torch.randn(1003, device="cuda"); g = torch.randn(1003, device="cuda").bfloat16()
    ref = acc + g.float(); ops.score_accumulate(acc, g)
    ok &= report("score_accumulate tail", torch.equal(acc, ref))
    for b in (64, 128, 256):
        a = torch.randn(1024, 768, device="cuda")
        v = a.reshape(1024 // b, b, 768 // b, b)
        refs = {"mean_abs": v.mean(dim=(1, 3)).abs(), "abs_mean": v.abs().mean(dim=(1, 3)),
                "L1": v.abs().sum(dim=(1, 3)), "L2": torch.sqrt(torch.sum(v.abs() ** 2, dim=(1, 3)))}
        for s, r in refs.items():
            out = ops.block_score_reduce(a, b, s)
            err = ((out - r).abs() / r.abs().clamp_min(1e-6)).max().item()
            ok &= report(f"block_score_reduce b={b} {s}", err < 2e-5 if s != "mean_abs" else
                         (out - r).abs().max().item() < 1e-6, f"relerr={err:.2e}")
        for dt in (torch.bfloat16, torch.float32):
            g = torch.randn(1024, 768, device="cuda").to(dt)
            sums = torch.zeros(1024 // b, 768 // b, device="cuda")
            ops.block_sum_accumulate(sums, g, b); ops.block_sum_accumulate(sums, g, b)
            r = 2 * g.float().reshape(1024 // b, b, 768 // b, b).sum(dim=(1, 3))
            ok &= report(f"block_sum_accumulate b={b} {dt}", (sums - r).abs().max().item() < 1e-2 * b / 64,
                         f"abserr={(sums - r).abs().max().item():.2e}")
            fin = ops.block_sum_finalize(sums, b)
            ok &= report(f"block_sum_finalize b={b}", torch.allclose(fin, (sums / (b * b)).abs()))
    x = torch.randn(3, 40, 512, device="cuda").bfloat16()
    acc = torch.zeros(40, 512, device="cuda")
    ops.act_score_accumulate(acc, x); ops.act_score_accumulate(acc, x)
    r = 2 * x.float().abs().sum(0)
    ok &= report("act_score_accumulate", torch.allclose(acc, r, rtol=1e-5, atol=1e-5))
    for s in ("mean_abs", "abs_mean", "L1", "L2"):
        out = ops.channel_score_reduce(acc, s)
        rr = {"mean_abs": r.abs().mean(0), "abs_mean": r.mean(0).abs(), "L1": r.abs().sum(0),
              "L2": torch.norm(r, p=2, dim=0)}[s]
        ok &= report(f"channel_score_reduce {s}", torch.allclose(out, rr, rtol=1e-4, atol=1e-4))
    return ok


def g_topk():
    import torch
    from sparse_matrix_tuning_b200 import ops
    torch.manual_seed(1)
    ok = True
    for n, k in ((1000, 17), (12288, 869), (98304, 2106), (5000, 5000), (40000, 20000), (1, 1), (300, 0)):
        sc = torch.randn(n, device="cuda").abs()
        sc[::7] = sc[0].item()  # plant ties
        rank = torch.randperm(n, device="cuda").to(torch.int32)
        inv = torch.empty_like(rank); inv[rank.long()] = torch.arange(n, device="cuda", dtype=torch.int32)
        idx, offs = ops.topk_blocks(sc, [0, n], [k], rank, inv)
        # reference: sort by (score, rank) descending
        key = list(zip(sc.tolist(), rank.tolist(), range(n)))
        key.sort(reverse=True)
        ref = [t[2] for t in key[:k]]
        ok &= report(f"topk n={n} k={k}", idx.tolist() == ref)
    # segmented, no rank (ties -> larger index first)
    sc = torch.randn(3000, device="cuda")
    offs = [0, 1000, 1000, 2500, 3000]
    ks = [10, 5, 2000, 7]
    idx, oo = ops.topk_blocks(sc, offs, ks)
    good = True
    l = sc.tolist()
    for s in range(4):
        seg = sorted([(l[i], i) for i in range(offs[s], offs[s + 1])], reverse=True)[: ks[s]]
        good &= idx[oo[s]:oo[s + 1]].tolist() == [t[1] for t in seg]
    ok &= report("topk segmented", good)
    return ok


def g_copy():
    import torch
    from sparse_matrix_tuning_b200 import ops
    torch.manual_seed(2)
    ok = True
    for b in (64, 128, 256):
        for dt in (torch.bfloat16, torch.float32):
            W1 = torch.randn(512, 768, device="cuda").to(dt)
            W2 = torch.randn(1024, 512, device="cuda").to(dt)
            ents = [(W1, 1, 2), (W1, 0, 0), (W2, 3, 1), (W2, 1, 1)]
            ents = [(w, r % (w.shape[0] // b), c % (w.shape[1] // b)) for w, r, c in ents]
            tab = ops.make_block_table(ents, "cuda")
            comp = torch.empty(len(ents) * b, b, device="cuda", dtype=dt)
            ops.block_gather(tab, len(ents), b, comp)
            ref = torch.cat([w[r * b:(r + 1) * b, c * b:(c + 1) * b] for w, r, c in ents])
            ok &= report(f"gather b={b} {dt}", torch.equal(comp, ref))
            comp2 = torch.randn_like(comp)
            W1c, W2c = W1.clone(), W2.clone()
            ops.block_scatter(tab, len(ents), b, comp2)
            for i, (w, r, c) in enumerate(ents):
                wc = W1c if w is W1 else W2c
                wc[r * b:(r + 1) * b, c * b:(c + 1) * b] = comp2[i * b:(i + 1) * b]
            ok &= report(f"scatter b={b} {dt}", torch.equal(W1, W1c) and torch.equal(W2, W2c))
    return ok


def g_adam():
    import torch
    from sparse_matrix_tuning_b200 import ops
    torch.manual_seed(3)
    ok = True
    b, n = 64, 6
    N = n * b * b
    W = torch.randn(256, 256, device="cuda").bfloat16()
    ents = [(W, i // 4, i % 4) for i in range(n)]
    tab = ops.make_block_table(ents, "cuda")
    master = torch.empty(N, device="cuda")
    ops.block_gather(tab, n, b, comp := torch.empty(n * b, b, device="cuda", dtype=torch.bfloat16))
    master.copy_(comp.float().flatten())
    m = torch.zeros(N, device="cuda"); v = torch.zeros(N, device="cuda")
    p_ref = master.clone().requires_grad_(True)
    opt = torch.optim.AdamW([p_ref], lr=1e-3, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.01)
    comp_out = torch.empty(n * b, b, device="cuda", dtype=torch.bfloat16)
    for step in range(1, 4):
        g = (torch.randn(N, device="cuda") * 0.1).bfloat16()
        sq = ops.grad_sqnorm(g)
        ref_sq = (g.float() ** 2).sum()
        ok &= report(f"sqnorm step {step}", abs(sq.item() - ref_sq.item()) / ref_sq.item() < 1e-5)
        ops.compact_adam(master, m, v, g, lr=1e-3, beta1=0.9, beta2=0.95, eps=1e-8, weight_decay=0.01,
                         step=step, grad_scale=1.0, sqnorm=sq, max_norm=1.0, compact_out=comp_out, table=tab,
                         n_blocks=n, block=b, w_dtype=torch.bfloat16)
        p_ref.grad = g.float().clone()
        torch.nn.utils.clip_grad_norm_([p_ref], 1.0)
        opt.step()
        err = (master - p_ref.detach()).abs().max().item()
        ok &= report(f"adam step {step} vs torch.AdamW", err < 2e-6, f"maxabs={err:.2e}")
    ok &= report("adam compact_out", torch.equal(comp_out.flatten(), master.bfloat16()))
    Wr = torch.cat([W[r * b:(r + 1) * b, c * b:(c + 1) * b] for _, r, c in ents])
    ok &= report("adam write-back", torch.equal(Wr, comp_out))
    return ok


def _gemm_case(b, T, n, dt, out_dt, accumulate=False, feat_in=1024, feat_out=768, seed=0):
    import torch
    from sparse_matrix_tuning_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.randn(T, feat_in, device="cuda", generator=g).to(dt)
    dy = torch.randn(T, feat_out, device="cuda", generator=g).to(dt)
    rc = [(int(torch.randint(0, feat_out // b, (1,), generator=torch.Generator().manual_seed(seed + i))),
           int(torch.randint(0, feat_in // b, (1,), generator=torch.Generator().manual_seed(100 + seed + i))))
          for i in range(n)]
    tab = ops.make_block_rc(rc, "cuda")
    out = None
    base = None
    if accumulate:
        base = torch.randn(n * b, b, device="cuda", generator=g).to(out_dt)
        out = base.clone()
    G = ops.block_grad_gemm(x, dy, tab, b, out=out, out_dtype=out_dt, accumulate=accumulate)
    torch.cuda.synchronize()
    ref = torch.cat([dy.float()[:, r * b:(r + 1) * b].t() @ x.float()[:, c * b:(c + 1) * b] for r, c in rc])
    if accumulate:
        ref = ref + base.float()
    scale = ref.abs().max().item()
    err = (G.float() - ref).abs().max().item() / scale
    tol = 1e-5 if out_dt == torch.float32 and dt == torch.float32 else (2e-5 if out_dt == torch.float32 else 8e-3)
    splits, ctas = ops.block_grad_gemm_plan(n, b, T, dt)
    return report(f"gemm b={b} T={T} n={n} {dt}->{out_dt} acc={accumulate} splits={splits}", err < tol,
                  f"relerr={err:.2e}")


def g_gemm_small():
    import torch
    ok = True
    for b in (256, 128, 64):
        ok &= _gemm_case(b, 64, 1, torch.bfloat16, torch.float32)
        ok &= _gemm_case(b, 256, 3, torch.bfloat16, torch.float32)
        ok &= _gemm_case(b, 200, 2, torch.bfloat16, torch.bfloat16)     # ragged T (TMA zero fill)
        ok &= _gemm_case(b, 256, 2, torch.float16, torch.float32)
        ok &= _gemm_case(b, 192, 2, torch.bfloat16, torch.bfloat16, accumulate=True)
    return ok


def g_gemm_split():
    import torch
    ok = True
    for b in (256, 128, 64):
        ok &= _gemm_case(b, 4096, 5, torch.bfloat16, torch.float32, feat_in=2048, feat_out=2048)
        ok &= _gemm_case(b, 5000, 3, torch.bfloat16, torch.bfloat16, feat_in=2048, feat_out=1024)
        ok &= _gemm_case(b, 4096, 2, torch.bfloat16, torch.bfloat16, accumulate=True)
        ok &= _gemm_case(b, 1024, 200, torch.bfloat16, torch.float32, feat_in=4096, feat_out=4096)
    return ok


def g_gemm_f32():
    import torch
    ok = True
    for b in (256, 128, 64):
        ok &= _gemm_case(b, 128, 3, torch.float32, torch.float32)
        ok &= _gemm_case(b, 77, 2, torch.float32, torch.float32, accumulate=True)
    return ok


def g_perf():
    import torch
    from sparse_matrix_tuning_b200 import ops
    ok = True
    # HBM kernels
    R, Cc = 16384, 16384
    acc = torch.zeros(R, Cc, device="cuda")
    g = torch.randn(R, Cc, device="cuda").bfloat16()
    ms = _time_cuda(lambda: ops.score_accumulate(acc, g))
    print(f"score_accumulate {R}x{Cc}: {ms:.3f} ms  {R * Cc * 10 / ms / 1e6:.0f} GB/s (10 B/elt)")
    ms = _time_cuda(lambda: ops.block_score_reduce(acc, 256, "mean_abs"))
    print(f"block_score_reduce b=256: {ms:.3f} ms  {R * Cc * 4 / ms / 1e6:.0f} GB/s")
    sums = torch.zeros(R // 256, Cc // 256, device="cuda")
    ms = _time_cuda(lambda: ops.block_sum_accumulate(sums, g, 256))
    print(f"block_sum_accumulate b=256 bf16: {ms:.3f} ms  {R * Cc * 2 / ms / 1e6:.0f} GB/s")
    del acc, g
    n, b = 869, 256
    N = n * b * b
    W = torch.zeros(4096, 4096 * 4, device="cuda", dtype=torch.bfloat16)
    ents = [(W, i // 64, i % 64) for i in range(n)]
    tab = ops.make_block_table(ents, "cuda")
    master = torch.zeros(N, device="cuda"); m = torch.zeros(N, device="cuda"); v = torch.zeros(N, device="cuda")
    gr = torch.randn(N, device="cuda").bfloat16()
    sq = ops.grad_sqnorm(gr)
    ms = _time_cuda(lambda: ops.compact_adam(master, m, v, gr, lr=1e-4, beta1=0.9, beta2=0.95, eps=1e-8,
                                             weight_decay=0.0, step=3, sqnorm=sq, max_norm=1.0, table=tab,
                                             n_blocks=n, block=b, w_dtype=torch.bfloat16))
    print(f"compact_adam 869 blocks: {ms:.3f} ms  {N * 28 / ms / 1e6:.0f} GB/s (28 B/elt)")
    ms = _time_cuda(lambda: ops.grad_sqnorm(gr, sq))
    print(f"grad_sqnorm: {ms:.3f} ms  {N * 2 / ms / 1e6:.0f} GB/s")
    del master, m, v, gr, W
    # GEMM
    for (b, T, n, fi, fo) in ((256, 16384, 13, 4096, 4096), (256, 8192, 9, 4096, 4096), (256, 16384, 148, 4096, 14336),
                              (256, 65536, 148, 4096, 4096), (128, 16384, 51, 4096, 4096), (64, 16384, 204, 4096, 4096)):
        x = torch.randn(T, fi, device="cuda").bfloat16()
        dy = torch.randn(T, fo, device="cuda").bfloat16()
        perm = torch.randperm((fo // b) * (fi // b))[:n]
        rc = [(int(p) // (fi // b), int(p) % (fi // b)) for p in perm]
        tab = ops.make_block_rc(rc, "cuda")
        out = torch.empty(n * b, b, device="cuda", dtype=torch.bfloat16)
        ms = _time_cuda(lambda: ops.block_grad_gemm(x, dy, tab, b, out=out))
        fl = 2.0 * b * b * T * n
        splits, ctas = ops.block_grad_gemm_plan(n, b, T, torch.bfloat16)
        print(f"block_grad_gemm b={b} T={T} n={n} splits={splits} ctas={ctas}: {ms * 1e3:.1f} us  {fl / ms / 1e9:.1f} TFLOP/s")
        del x, dy
    return ok


def main():
    if len(sys.argv) > 1 and sys.argv[1].startswith("--run="):
        name = sys.argv[1][6:]
        ok = globals()["g_" + name]()
        sys.exit(0 if ok else 1)
    groups = sys.argv[1:] or GROUPS
    summary = {}
    for gname in groups:
        t = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), f"--run={gname}"], timeout=600)
            summary[gname] = r.returncode
        except subprocess.TimeoutExpired:
            summary[gname] = "timeout"
        print(f"== group {gname}: rc={summary[gname]} ({time.time() - t:.1f}s)", flush=True)
    print("SUMMARY", summary)
    sys.exit(0 if all(v == 0 for v in summary.values()) else 1)


if __name__ == "__main__":
    main()
