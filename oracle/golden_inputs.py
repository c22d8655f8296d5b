"""Seeded synthetic inputs shared by oracle/gen_golden.py (which feeds them to the real reference) and by
the tests (which feed them to the oracle and to the CUDA path).  TEST INFRASTRUCTURE ONLY.

Inputs are produced by torch CPU generators with fixed seeds (bit-reproducible for a fixed torch build, which
the build container and the GPU box share); each fixture also records a SHA-256 of the input bytes so that any
drift is detected instead of silently comparing different problems.
"""
from __future__ import annotations

import hashlib

import torch

ATTN_SHAPES_SMALL = {"q_proj": [512, 512], "k_proj": [256, 512], "v_proj": [256, 512]}
ATTN_SHAPES_MED = {"q_proj": [1024, 1024], "k_proj": [256, 1024], "v_proj": [256, 1024]}
MLP_SHAPES = {"gate_proj": [1536, 512], "up_proj": [1536, 512], "down_proj": [512, 1536]}


def _spec(name, shapes, layers, n, sel="no_restriction", calc="mean_abs", seed=0, kind="gauss"):
    return {"name": name, "shapes": shapes, "layers": layers, "n": n, "selection_strategy": sel,
            "calculate_strategy": calc, "seed": seed, "kind": kind}


SELECTION_SPECS = [
    _spec("attn_small_mean_abs", ATTN_SHAPES_SMALL, [0, 1], 5, seed=11),
    _spec("attn_small_abs_mean", ATTN_SHAPES_SMALL, [0, 1], 5, calc="abs_mean", seed=12),
    _spec("attn_small_L1", ATTN_SHAPES_SMALL, [0, 1], 7, calc="L1", seed=13),
    _spec("attn_small_L2", ATTN_SHAPES_SMALL, [0, 1], 7, calc="L2", seed=14),
    _spec("attn_med_mean_abs", ATTN_SHAPES_MED, [0, 1, 2, 3], 40, seed=15),
    _spec("mlp_L2", MLP_SHAPES, [0, 1], 10, calc="L2", seed=16),
    _spec("attn_norm_dist", ATTN_SHAPES_SMALL, [0, 1], 2, sel="norm_dist", seed=17),
    _spec("mlp_norm_dist_abs_mean", MLP_SHAPES, [3], 4, sel="norm_dist", calc="abs_mean", seed=18),
    # exact score ties: every block is a constant from a tiny set, so the tuple-order tie rule decides
    _spec("ties_mean_abs", ATTN_SHAPES_SMALL, [0, 1], 9, seed=19, kind="ties"),
    _spec("ties_layers_9_10_11", ATTN_SHAPES_SMALL, [9, 10, 11], 13, seed=20, kind="ties"),
    _spec("ties_L1", MLP_SHAPES, [2, 10], 11, calc="L1", seed=21, kind="ties"),
    _spec("all_equal", ATTN_SHAPES_SMALL, [0, 1], 6, seed=22, kind="ones"),
    _spec("n_exceeds_blocks", ATTN_SHAPES_SMALL, [0], 100, seed=23),
    _spec("n_zero", ATTN_SHAPES_SMALL, [0], 0, seed=24),
    _spec("n_one", ATTN_SHAPES_MED, [0, 1], 1, seed=25),
    _spec("sign_cancel", ATTN_SHAPES_SMALL, [0, 1], 4, seed=26, kind="cancel"),
]


def make_selection_inputs(spec, block: int = 256):
    """{(module, layer): fp32 CPU tensor}, targeted_module_dims — dict order = (layer, module) nesting as the
    driver's capture loop produces it (named_parameters order)."""
    g = torch.Generator().manual_seed(spec["seed"])
    grads = {}
    for layer in spec["layers"]:
        for mod, (r, c) in spec["shapes"].items():
            if spec["kind"] == "gauss":
                t = torch.randn(r, c, generator=g)
            elif spec["kind"] == "ones":
                t = torch.ones(r, c)
            elif spec["kind"] == "ties":
                vals = torch.randint(1, 4, (r // block, c // block), generator=g).float()
                sign = torch.randint(0, 2, (r // block, c // block), generator=g).float() * 2 - 1
                t = (vals * sign).repeat_interleave(block, 0).repeat_interleave(block, 1).contiguous()
            elif spec["kind"] == "cancel":
                # large |g| with zero block mean next to small |g| with non-zero mean: mean_abs must prefer the latter
                t = torch.randn(r, c, generator=g)
                t[:block, :block] = 50.0
                t[:block, :block // 2] = -50.0
            else:
                raise ValueError(spec["kind"])
            grads[(mod, layer)] = t
    return grads, {k: list(v) for k, v in spec["shapes"].items()}


CHANNEL_SPECS = [
    {"name": "planted", "n": 100, "selection_strategy": "no_restriction", "calculate_strategy": "mean_abs", "seed": 0,
     "kind": "planted"},
    {"name": "gauss_L2", "n": 37, "selection_strategy": "no_restriction", "calculate_strategy": "L2", "seed": 31,
     "kind": "gauss"},
    {"name": "gauss_L1_norm_dist", "n": 5, "selection_strategy": "norm_dist", "calculate_strategy": "L1", "seed": 32,
     "kind": "gauss"},
    {"name": "gauss_abs_mean", "n": 16, "selection_strategy": "no_restriction", "calculate_strategy": "abs_mean",
     "seed": 33, "kind": "gauss"},
]


def make_channel_inputs(spec):
    """Scaled-down restatement of the planted-pattern example in the reference's smt_helper.py:323-334."""
    if spec["kind"] == "planted":
        act = {("gate_proj", 1): torch.zeros(3, 96, 512), ("up_proj", 1): torch.zeros(3, 96, 512),
               ("down_proj", 2): torch.ones(3, 64, 768)}
        act[("gate_proj", 1)][:, :, 0:64] = 1.0
        act[("gate_proj", 1)][:, :, 0:4] = 10.0
        act[("up_proj", 1)][:, :, 3:6] = 100.0
        act[("down_proj", 2)][:, :, 3:6] = 100.0
        return act
    g = torch.Generator().manual_seed(spec["seed"])
    return {("q_proj", 0): torch.randn(2, 48, 256, generator=g).abs(),
            ("k_proj", 0): torch.randn(2, 48, 256, generator=g).abs() * 1.5,
            ("down_proj", 1): torch.randn(2, 48, 512, generator=g).abs()}


LINEARZ_SPECS = [
    {"name": "fp32_b256", "dtype": "float32", "B": 2, "S": 32, "in": 512, "out": 512, "block": 256,
     "index_list": [(0, 1), (1, 0), (1, 1)], "seed": 41},
    {"name": "bf16_b256_rect", "dtype": "bfloat16", "B": 2, "S": 64, "in": 512, "out": 768, "block": 256,
     "index_list": [(2, 1), (0, 0)], "seed": 42},
    {"name": "bf16_b256_B4", "dtype": "bfloat16", "B": 4, "S": 40, "in": 768, "out": 256, "block": 256,
     "index_list": [(0, 2), (0, 0), (0, 1)], "seed": 43},
    {"name": "bf16_b64", "dtype": "bfloat16", "B": 2, "S": 48, "in": 256, "out": 128, "block": 64,
     "index_list": [(1, 3), (0, 0), (1, 0), (0, 2), (1, 1)], "seed": 44},
    {"name": "fp32_b128_B1", "dtype": "float32", "B": 1, "S": 50, "in": 256, "out": 384, "block": 128,
     "index_list": [(2, 1), (0, 0)], "seed": 45},
]


# linearChannel (smt.py:220-296) runs in the reference only for SQUARE weights (it copies weight rows into an
# [n, in_features] parameter but produces an [n, out_features] gradient)
LINEARCHANNEL_SPECS = [
    {"name": "fp32_sq", "dtype": "float32", "B": 2, "S": 24, "in": 256, "out": 256,
     "index_list": [200, 3, 77, 78, 255, 0], "seed": 51},
    {"name": "bf16_sq_B4", "dtype": "bfloat16", "B": 4, "S": 40, "in": 512, "out": 512,
     "index_list": [511, 17, 300, 301, 302, 5, 64, 128, 256, 9, 100], "seed": 52},
    {"name": "bf16_sq_B1", "dtype": "bfloat16", "B": 1, "S": 96, "in": 384, "out": 384,
     "index_list": [1], "seed": 53},
]


def make_linearchannel_inputs(spec):
    g = torch.Generator().manual_seed(spec["seed"])
    dt = getattr(torch, spec["dtype"])
    x = torch.randn(spec["B"], spec["S"], spec["in"], generator=g).to(dt)
    dy = torch.randn(spec["B"], spec["S"], spec["out"], generator=g).to(dt)
    w = (torch.randn(spec["out"], spec["in"], generator=g) * 0.02).to(dt)
    return x, dy, w, [int(i) for i in spec["index_list"]]


def make_linearz_inputs(spec):
    g = torch.Generator().manual_seed(spec["seed"])
    dt = getattr(torch, spec["dtype"])
    x = torch.randn(spec["B"], spec["S"], spec["in"], generator=g).to(dt)
    dy = torch.randn(spec["B"], spec["S"], spec["out"], generator=g).to(dt)
    w = (torch.randn(spec["out"], spec["in"], generator=g) * 0.02).to(dt)
    return x, dy, w, [tuple(t) for t in spec["index_list"]]


CONFIG1 = {"vocab_size": 32000, "hidden_size": 512, "intermediate_size": 1536, "num_hidden_layers": 2,
           "num_attention_heads": 8, "num_key_value_heads": 4, "max_position_embeddings": 512,
           "batch": 2, "seq": 64, "seed": 1234, "attn_ratio": 0.01, "warmup_steps": 2, "sparse_steps": 4,
           "smt_lr": 1e-4}


def make_config1(device="cpu"):
    """2-layer random-init LlamaForCausalLM (fp32, built on the CPU with seed 1234 = the reference's default seed,
    fine_tune.py:964-967) and six seeded [2, 64] token batches."""
    from transformers import LlamaConfig, LlamaForCausalLM
    c = CONFIG1
    torch.manual_seed(c["seed"])
    cfg = LlamaConfig(vocab_size=c["vocab_size"], hidden_size=c["hidden_size"],
                      intermediate_size=c["intermediate_size"], num_hidden_layers=c["num_hidden_layers"],
                      num_attention_heads=c["num_attention_heads"], num_key_value_heads=c["num_key_value_heads"],
                      max_position_embeddings=c["max_position_embeddings"], attn_implementation="eager")
    model = LlamaForCausalLM(cfg).float()
    g = torch.Generator().manual_seed(c["seed"] + 1)
    batches = [torch.randint(0, c["vocab_size"], (c["batch"], c["seq"]), generator=g)
               for _ in range(c["warmup_steps"] + c["sparse_steps"])]
    if device != "cpu":
        model = model.to(device)
        batches = [b.to(device) for b in batches]
    return model, batches


def tensor_dict_sha(d) -> str:
    h = hashlib.sha256()
    for k in d:
        h.update(repr(k).encode())
        h.update(d[k].detach().contiguous().cpu().view(torch.uint8).numpy().tobytes())
    return h.hexdigest()
