#!/usr/bin/env python
"""bench.py — headline benchmark of the SMT hot path on B200.

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's algorithm on the box's host cores

Workload (BASELINE.json metric "tokens/sec/GPU (LLaMA-3-8B SMT 0.71%)", configs[2]/[3]): a random-init LLaMA-3-8B
(`LlamaForCausalLM`, bf16) whose q/k/v projections own 869 selected 256x256 blocks (0.71 % of the 122 528 blocks the
reference's budget counts, fine_tune.py:231-239).  One STEP = forward + backward (the selected-block gradients come
from the tcgen05 block-gradient GEMM) + [N>1: one all-reduce of the flat compact-gradient buffer] + one fused
compact-Adam/clip/write-back launch, on synthetic tokens (uniform ids, labels = ids).  Before the timed steps the
script runs the warm-up part of the path once (on-device block-score accumulation over one backward pass, block-score
finalize, exact top-k, freeze, convert) and reports its time separately.

One JSON line on stdout (rank 0).  `value` = tokens/s over all ranks with inputs resident in HBM; `e2e` = the same
step with the token ids copied from pinned host memory and the loss read back every step; `roofline` describes the
dominant SMT kernel (block-gradient GEMM) from CUDA-event timings taken live inside the timed region;
`cpu_baseline` times the oracle port of the reference's path on the host cores (bounded sample).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LLAMA3_8B = dict(vocab_size=128256, hidden_size=4096, intermediate_size=14336, num_hidden_layers=32,
                 num_attention_heads=32, num_key_value_heads=8, max_position_embeddings=8192, rope_theta=500000.0,
                 rms_norm_eps=1e-5, tie_word_embeddings=False)
ATTN_RATIO = 0.0071       # "SMT 0.71 %": int(0.0071 * 122528) = 869 blocks  
BLOCK = 256


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--batch", type=int, default=16, help="sequences per GPU")
    ap.add_argument("--seq", type=int, default=512)
    ap.add_argument("--layers", type=int, default=32, help="(debug only) fewer layers => config.workload says so")
    ap.add_argument("--no-ckpt", action="store_true", help="disable gradient checkpointing (reference: on, fine_tune.py:192)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seq", type=int, default=512, help="tokens of the bounded CPU sample")
    ap.add_argument("--attn-ratio", type=float, default=ATTN_RATIO, help="fine_tune.py --downsample_attention_blocks_ratio")
    ap.add_argument("--mlp-ratio", type=float, default=0.0, help="fine_tune.py --downsample_mlp_blocks_ratio (0 = off)")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra no-checkpointing measurement")
    ap.add_argument("--no-group", action="store_true", help="launch the block-gradient GEMM per module instead of grouped")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------------------------------------------

class ClockSampler:
    """nvidia-smi sampled every 200 ms while the timed region runs (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                                  ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's path restated by oracle/ (the reference itself is pure Python and is not on the GPU box)
# ---------------------------------------------------------------------------------------------------------------

def cpu_reference_arm(steps: int, warmup: int, seq: int):
    """Bounded sample: ONE LLaMA-3-8B decoder layer (of 32) in bf16 on the host cores, q/k/v converted exactly as the
    reference converts them (27 = round(869/32) selected 256x256 blocks), batch 1 x `seq` tokens: forward (with the
    per-forward scatter loop), backward (per-block bmm + sum + copy loop, smt.py:386-404) and a clipped AdamW step
    on the compact parameters.  tokens/s is extrapolated as seq / (32 * t_layer): embeddings, final norm, lm_head and
    the loss are left out, which favours the CPU arm."""
    import torch
    from transformers import LlamaConfig
    from transformers.models.llama.modeling_llama import LlamaDecoderLayer, LlamaRotaryEmbedding
    from oracle import smt_oracle as O

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(1234)
    cfg = LlamaConfig(**{**LLAMA3_8B, "num_hidden_layers": 1}, attn_implementation="sdpa")
    layer = LlamaDecoderLayer(cfg, layer_idx=0).to(torch.bfloat16)
    rope = LlamaRotaryEmbedding(cfg)
    for p in layer.parameters():
        p.requires_grad = False
    g = torch.Generator().manual_seed(7)
    per_module = {"q_proj": 15, "k_proj": 6, "v_proj": 6}            # 27 blocks/layer, k/v (GQA) are 4 x 16 blocks
    sparse = []
    for name, n in per_module.items():
        lin = getattr(layer.self_attn, name)
        rows, cols = lin.weight.shape[0] // BLOCK, lin.weight.shape[1] // BLOCK
        perm = torch.randperm(rows * cols, generator=g)[:n]
        idx = [(int(p) // cols, int(p) % cols) for p in perm]
        mod = O.OracleSparseLinear(lin.weight, idx, BLOCK)
        setattr(layer.self_attn, name, mod)
        sparse.append(mod)
    params = [m.selected_weight for m in sparse]
    opt = torch.optim.AdamW(params, lr=1e-4, betas=(0.9, 0.95), eps=1e-8, weight_decay=0.0)
    x = torch.randn(1, seq, cfg.hidden_size, generator=g).to(torch.bfloat16).requires_grad_(True)
    pos = torch.arange(seq).unsqueeze(0)
    cos_sin = rope(x, pos)

    def one_step():
        opt.zero_grad()
        out = layer(x, position_embeddings=cos_sin, position_ids=pos, attention_mask=None)
        out = out[0] if isinstance(out, tuple) else out
        out.float().pow(2).mean().backward()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()

    for _ in range(warmup):
        one_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    tokens_per_s = seq / (dt * LLAMA3_8B["num_hidden_layers"])
    sample = (f"1 of 32 LLaMA-3-8B decoder layers, bf16, batch 1 x {seq} tokens, 27 q/k/v blocks, oracle port of "
              f"smt.py forward/backward + clipped AdamW; tokens/s = {seq}/(32*t_layer); {steps} steps after {warmup} warm-up")
    return tokens_per_s, dt * 1e3, cores, sample


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    tps, ms, cores, sample = cpu_reference_arm(args.steps, args.warmup, args.cpu_seq)
    line = {"impl": "reference", "metric": "tokens/sec/GPU (LLaMA-3-8B SMT 0.71%)", "value": tps, "unit": "tokens/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms * LLAMA3_8B["num_hidden_layers"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"LLaMA-3-8B SMT 0.71% q/k/v gradient-based selection, bf16, seq {args.seq} x batch "
                                   f"{args.batch} per GPU",
                       "note": "same workload as the GPU arm, measured on a bounded sample (see cpu_baseline.sample); the "
                               "reference is pure Python/PyTorch and is not present on the GPU box, so this arm runs the "
                               "oracle port of its hot path (oracle/smt_oracle.py) on all host cores"},
            "cpu_baseline": {"value": tps, "unit": "tokens/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": tps, "unit": "tokens/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------

def build_model(args, device):
    import torch
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = LlamaConfig(**{**LLAMA3_8B, "num_hidden_layers": args.layers}, attn_implementation="sdpa")
    torch.set_default_dtype(torch.bfloat16)
    try:
        with torch.device(device):
            model = LlamaForCausalLM(cfg)
    finally:
        torch.set_default_dtype(torch.float32)
    model.config.use_cache = False
    return model


def run_ours(args):
    import torch
    import torch.distributed as dist

    from sparse_matrix_tuning_b200 import _lib, dp, ops
    from sparse_matrix_tuning_b200.optim import SMTAdam
    from sparse_matrix_tuning_b200.smt import smt as M, smt_helper as H
    from sparse_matrix_tuning_b200.warmup import WarmupGradAccumulator

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    _lib.load()                                                   # fail loudly if the extension is missing

    torch.manual_seed(1234)                                       # identical weights on every rank
    model = build_model(args, device)
    vocab = LLAMA3_8B["vocab_size"]
    gen = torch.Generator(device="cpu").manual_seed(1234 + rank)  # per-rank data (weak scaling)
    B, S = args.batch, args.seq
    n_batches = args.warmup + args.steps + 1
    host_ids = [torch.randint(0, vocab, (B, S), generator=gen).pin_memory() for _ in range(n_batches)]
    dev_ids = [t.to(device) for t in host_ids]

    # ---- warm-up part of the path (once, timed separately): capture -> scores -> top-k -> freeze -> convert -------
    named = list(model.named_parameters())
    dims = {}
    for name, p in named:                                         # fine_tune.py:221-228
        if "weight" in name:
            for t in ("gate_proj", "up_proj", "down_proj", "q_proj", "k_proj", "v_proj"):
                if t in name and t not in dims:
                    dims[t] = [p.shape[0], p.shape[1]]
                    break
    total_blocks = sum(p.shape[0] / BLOCK * p.shape[1] / BLOCK for _n, p in named if p.ndim == 2)   # fine_tune.py:231-234
    n_attn = int(args.attn_ratio * total_blocks)                  # fine_tune.py:236
    n_mlp = int(args.mlp_ratio * total_blocks)                    # fine_tune.py:239 (0 = MLP not selected, the default)
    for name, p in named:                                         # capture needs q/k/v (+ MLP) weight gradients only
        p.requires_grad = (("self_attn" in name) and any(k in name for k in ("q_proj", "k_proj", "v_proj"))) or \
                          (n_mlp > 0 and "mlp" in name)
    if not args.no_ckpt:
        # non-reentrant checkpointing keeps the whole backward in ONE autograd graph task, so the block-gradient GEMMs
        # of all modules can be deferred to a single grouped launch at the end of the pass
        model.gradient_checkpointing_enable(gradient_checkpointing_kwargs={"use_reentrant": False})
        model.enable_input_require_grads()
    model.train()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    # two accumulators, as the driver keeps `attention_warmup_grads` and `warmup_grads` apart (fine_tune.py:723-765)
    acc = WarmupGradAccumulator(block=BLOCK, mode="block_sum", mlp=False, attention=True)
    acc_mlp = WarmupGradAccumulator(block=BLOCK, mode="block_sum", mlp=True, attention=False) if n_mlp > 0 else None
    out = model(input_ids=dev_ids[-1], labels=dev_ids[-1], use_cache=False)
    out.loss.backward()
    acc.accumulate(model.named_parameters())
    dp.allreduce_block_sums(acc)                                  # scores of the DP-mean gradient on every rank
    if acc_mlp is not None:
        acc_mlp.accumulate(model.named_parameters())
        dp.allreduce_block_sums(acc_mlp)
    torch.cuda.synchronize()
    t_capture = time.perf_counter() - t0
    t0 = time.perf_counter()
    keys, scores = acc.scores("mean_abs")
    sel = H.select_submatrix_from_scores(keys, scores, n_attn, "no_restriction")       # fine_tune.py:306-313
    sel_mlp = {}
    if acc_mlp is not None:
        keys_m, scores_m = acc_mlp.scores("mean_abs")
        sel_mlp = H.select_submatrix_from_scores(keys_m, scores_m, n_mlp, "no_restriction")   # fine_tune.py:319-327
    torch.cuda.synchronize()
    t_select = time.perf_counter() - t0
    dp.assert_same_selection(sel)
    dp.assert_same_selection(sel_mlp)
    model.zero_grad(set_to_none=True)
    del acc, acc_mlp, out
    model = M.freeze_unselected_matrix_layer(model, sel_mlp, sel)
    model = M.convert_linear_layer_to_matrix_sparsity(model, sel_mlp, sel)
    sel = {**sel, **sel_mlp}                                      # below: bookkeeping over every converted module
    groups = M.get_optimizer_sparse_grouped_parameters(model, 0.0, 1e-4)
    opt = SMTAdam(groups, lr=1e-4, betas=(0.9, 0.95), max_grad_norm=1.0)
    n_blocks = sum(len(v) for v in sel.values())
    trainable = opt.trainable_elements()
    M.set_grouped_backward(not args.no_group)
    torch.cuda.empty_cache()

    def step(ids):
        out = model(input_ids=ids, labels=ids, use_cache=False)
        out.loss.backward()
        works = dp.allreduce_compact_grads(opt, async_op=True)
        for w in works:
            w.wait()
        opt.step()
        opt.zero_grad()
        return out.loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(dev_ids[i])
    # ---- timed region 1: inputs resident in HBM ------------------------------------------------------------------
    ops.enable_timing("block_grad_gemm")
    ops.enable_timing("compact_adam")
    launches0 = ops.LAUNCHES["total"]
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        loss = step(dev_ids[args.warmup + i])
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    launches = ops.LAUNCHES["total"] - launches0
    gemm_t = ops.collect_timing("block_grad_gemm")
    adam_t = ops.collect_timing("compact_adam")
    ops.enable_timing("block_grad_gemm", False)
    ops.enable_timing("compact_adam", False)
    # ---- timed region 2: end to end (pinned host ids in, loss value out, every step) ------------------------------
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    last = None
    for i in range(args.steps):
        ids = host_ids[args.warmup + i].to(device, non_blocking=True)
        last = step(ids).item()
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    # ---- extra (reported, not the headline): the same step WITHOUT activation recomputation ------------------------
    ms_nockpt = None
    if not args.no_ckpt and not args.no_extra:
        model.gradient_checkpointing_disable()
        for i in range(2):
            step(dev_ids[i])
        barrier()
        e4, e5 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e4.record()
        for i in range(args.steps):
            step(dev_ids[args.warmup + i])
        e5.record()
        barrier()
        ms_nockpt = e4.elapsed_time(e5)
    if world > 1:
        t = torch.tensor([ms_total, ms_e2e, ms_nockpt or 0.0], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total, ms_e2e, ms_nockpt = t.tolist()
        ms_nockpt = ms_nockpt or None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    tokens = B * S * world * args.steps
    value = tokens / (ms_total / 1e3)
    e2e_value = tokens / (ms_e2e / 1e3)
    # ---- roofline of the dominant SMT kernel: block-gradient GEMM ---------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops_sustained") or 1400.0         # sustained: the kernel is timed inside a long step
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained (of measured)" if peaks else "fallback 1.4 PFLOP/s sustained (of fallback)"
    T = B * S
    gemm_ms = sum(ms for ms, _ in gemm_t)
    gemm_flops = sum(2.0 * tag[1] * tag[1] * tag[2] * tag[0] for _ms, tag in gemm_t)
    n_gemm = len(gemm_t)
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    # minimum HBM bytes of one step's block gradients: every distinct x / dy strip once + the bf16 outputs
    def x_group(kind):                                            # modules that read the same input activation
        return "attn" if kind in ("q_proj", "k_proj", "v_proj") else ("mlp_in" if kind in ("gate_proj", "up_proj") else kind)
    x_strips = {(x_group(k[0]), k[1], c) for k, idx in sel.items() for _r, c in idx}
    dy_strips = {(k, r) for k, idx in sel.items() for r, _c in idx}
    gemm_min_bytes = 2.0 * T * BLOCK * (len(x_strips) + len(dy_strips)) + 2.0 * n_blocks * BLOCK * BLOCK
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r01_bench_gemm_traffic.json")))
        if args.layers == 32 and not args.no_group and world == 1 and n_mlp == 0 and args.attn_ratio == ATTN_RATIO \
                and os.environ.get("SMT_GEMM_2SM") != "0":
            traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
    except Exception:
        pass
    adam_ms = statistics.mean(ms for ms, _ in adam_t) if adam_t else None
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    gemm_kernel = "block_grad_umma_kernel<256, 2, grouped> (single-CTA tiles)" if os.environ.get("SMT_GEMM_2SM") == "0" \
        else "block_grad_umma_2sm_kernel (cta_group::2, two 256x256 blocks per SM pair)"
    roofline = {"kernel": gemm_kernel, "bound": "tensor", "achieved": achieved,
                "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": traffic,
                "traffic_note": "dram read+write bytes per launch from one ncu --set full capture of this workload "
                                "(profiles/r01_bench_gemm_traffic.json); algorithmic minimum in min_hbm_bytes_per_launch",
                "min_hbm_bytes_per_launch": gemm_min_bytes * args.steps / max(n_gemm, 1),
                "frac_of_burst_peak": achieved / peaks.get("bf16_tflops", 1599.5),
                "peak_source": peak_src, "launches": n_gemm, "avg_launch_us": gemm_ms * 1e3 / max(n_gemm, 1),
                "flops_per_launch": gemm_flops / max(n_gemm, 1), "share_of_step": gemm_ms / ms_total,
                "also": {"compact_adam": {"bound": "hbm", "avg_ms": adam_ms,
                                          "achieved_gbs": (trainable * 30 / (adam_ms * 1e-3) / 1e9) if adam_ms else None,
                                          "peak_gbs": hbm_peak, "bytes_per_elem": 30,
                                          "frac": (trainable * 30 / (adam_ms * 1e-3) / 1e9 / hbm_peak) if adam_ms else None}}}
    line = {"metric": "tokens/sec/GPU (LLaMA-3-8B SMT 0.71%)", "value": value, "unit": "tokens/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": (f"LLaMA-3-8B SMT {100 * (args.attn_ratio + args.mlp_ratio):.2f}% "
                                    + ("q/k/v" if n_mlp == 0 else "q/k/v + MLP")
                                    + f" gradient-based selection, bf16, seq {S} x batch {B} per GPU")
                                   + ("" if args.layers == 32 else f" [DEBUG: {args.layers} layers only]"),
                       "selected_blocks": n_blocks, "block": BLOCK, "trainable_elements": trainable,
                       "modules_with_blocks": len(sel), "grouped_block_grad_launch": not args.no_group,
                       "grouped_launch_shape": dict(ops.LAST_GROUP),
                       "total_blocks_budget_base": total_blocks, "gradient_checkpointing": not args.no_ckpt,
                       "parallelism": f"dp{world}", "tokens_per_step_per_gpu": T,
                       "l2": "inputs larger than L2 (16 GB of weights streamed per step); no explicit flush",
                       "per_gpu_value": value / world, "loss_last": last,
                       "tokens_per_s_without_checkpointing": (tokens / (ms_nockpt / 1e3)) if ms_nockpt else None,
                       # capture = the process's FIRST forward + backward (cold: library init, allocator growth) + block-sum accumulation
                       "warmup_path_ms": {"capture_first_fwd_bwd_cold": t_capture * 1e3, "scores_topk": t_select * 1e3}},
            "e2e": {"value": e2e_value, "unit": "tokens/s", "h2d_bytes_per_step": B * S * 8, "d2h_bytes_per_step": 4,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline}
    if world == 1 and not args.no_cpu_baseline:
        try:
            tps, ms, cores, sample = cpu_reference_arm(steps=3, warmup=1, seq=args.cpu_seq)
            line["cpu_baseline"] = {"value": tps, "unit": "tokens/s", "cores": cores, "kind": "port", "sample": sample}
        except Exception as e:  # keep the GPU result even if the CPU leg cannot run
            line["cpu_baseline"] = {"value": None, "unit": "tokens/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"failed: {type(e).__name__}: {e}"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
