"""Generates the committed golden vectors under tests/golden/ by running the UNMODIFIED reference
(`/root/reference/deepspeed/smt/*.py`, imported through oracle/ref_shim.py) on seeded synthetic inputs.

TEST INFRASTRUCTURE ONLY.  Run in the build container (the reference does not exist on the GPU box):

    python oracle/gen_golden.py

Every fixture stores the reference's OUTPUTS; large inputs are re-created from the recorded seed with
`golden_inputs.py` helpers (shared with the tests) and guarded by a SHA-256 of the input bytes, small inputs are
stored verbatim.
"""
from __future__ import annotations

import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import golden_inputs as GI  # noqa: E402
from oracle.ref_shim import load_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def gen_selection(H):
    cases = []
    for spec in GI.SELECTION_SPECS:
        grads, dims = GI.make_selection_inputs(spec)
        kw = dict(selection_strategy=spec["selection_strategy"], calculate_strategy=spec["calculate_strategy"])
        try:
            sel = H.select_submatrix_based_on_grads(grads, dims, spec["n"], **kw)
        except Exception as e:  # e.g. n == 0 dies with UnboundLocalError at smt_helper.py:142
            cases.append({"spec": spec, "input_sha": GI.tensor_dict_sha(grads), "raises": type(e).__name__})
            continue
        scores = {}
        for key, g in grads.items():
            r = g.reshape(int(dims[key[0]][0] / 256), 256, int(dims[key[0]][1] / 256), 256)
            fn = {"mean_abs": H.mean_abs, "abs_mean": H.abs_mean_, "L1": H.L1_norm, "L2": H.L2_norm}[spec["calculate_strategy"]]
            scores[key] = fn(r).clone()
        cases.append({"spec": spec, "input_sha": GI.tensor_dict_sha(grads),
                      "selection": [(k, list(v)) for k, v in sel.items()], "scores": scores})
    torch.save(cases, os.path.join(OUT, "selection_cases.pt"))
    print(f"selection: {len(cases)} cases")


def gen_channels(H):
    cases = []
    for spec in GI.CHANNEL_SPECS:
        act = GI.make_channel_inputs(spec)
        sel = H.select_channel_based_on_activation(act, n=spec["n"], selection_strategy=spec["selection_strategy"],
                                                   calculate_strategy=spec["calculate_strategy"])
        cases.append({"spec": spec, "input_sha": GI.tensor_dict_sha(act),
                      "selection": [(k, list(v)) for k, v in sel.items()]})
    torch.save(cases, os.path.join(OUT, "channel_cases.pt"))
    print(f"channels: {len(cases)} cases")


def gen_linearz(S):
    cases = []
    for spec in GI.LINEARZ_SPECS:
        x, dy, w, index_list = GI.make_linearz_inputs(spec)
        S.Block_dimension = spec["block"]
        try:
            layer = S.LinearLayer_MatrixSparsity(torch.nn.Parameter(w.clone()), bias=None, index_list=index_list)
            selected0 = layer.selected_weight.detach().clone()
            # perturb the compact copy so that the forward's scatter is observable
            with torch.no_grad():
                layer.selected_weight.mul_(0.5)
            xin = x.clone().requires_grad_(True)
            y = layer(xin)
            y.backward(dy)
            cases.append({"spec": spec, "x": x, "dy": dy, "w": w, "index_list": index_list,
                          "selected0": selected0, "y": y.detach().clone(),
                          "w_after_sha": GI.tensor_dict_sha({"w": layer.weight.detach()}),
                          "grad_weight": layer.selected_weight.grad.detach().clone(),
                          "grad_input": xin.grad.detach().clone()})
        finally:
            S.Block_dimension = 256
    torch.save(cases, os.path.join(OUT, "linearz_cases.pt"))
    print(f"linearZ: {len(cases)} cases")


def gen_linearchannel(S):
    """The reference's channel layer on square weights: y, grad_input and the [n, out] channel gradient."""
    cases = []
    for spec in GI.LINEARCHANNEL_SPECS:
        x, dy, w, index_list = GI.make_linearchannel_inputs(spec)
        layer = S.LinearLayer_ChannelSparsity(torch.nn.Parameter(w.clone()), bias=None, index_list=index_list)
        xin = x.clone().requires_grad_(True)
        y = layer(xin)
        y.backward(dy)
        cases.append({"spec": spec, "x": x, "dy": dy, "w": w, "index_list": index_list,
                      "y": y.detach().clone(), "grad_weight": layer.selected_weight.grad.detach().clone(),
                      "grad_input": xin.grad.detach().clone()})
    torch.save(cases, os.path.join(OUT, "linearchannel_cases.pt"))
    print(f"linearChannel: {len(cases)} cases")


def gen_config1(S, H):
    """BASELINE config 1: 2-layer random-init LLaMA (hidden 512), 256x256 blocks, 1 % q/k/v, fp32 CPU.
    Flow (restating fine_tune.py): budget :231-239, two warm-up backward passes captured as :716-768 (no
    optimizer step: warm-up updates are the caller's full fine-tuning, not the SMT path), selection :306-313,
    freeze :334, convert :342, param groups :347, AdamW(betas=(0.9,0.95)) as FusedAdam stand-in :352-363 with
    global-norm clip 1.0 (deepspeed_helpers.py:87), four sparse steps."""
    model, batches = GI.make_config1()
    dims = {}
    for name, p in model.named_parameters():                     # fine_tune.py:221-228
        if "weight" in name:
            for t in ("gate_proj", "up_proj", "down_proj", "q_proj", "k_proj", "v_proj"):
                if t in name and t not in dims:
                    dims[t] = [p.shape[0], p.shape[1]]
                    break
    total = 0
    for _n, p in model.named_parameters():                       # fine_tune.py:231-234
        if p.ndim == 2:
            total += p.shape[0] / 256 * p.shape[1] / 256
    n_attn = int(GI.CONFIG1["attn_ratio"] * total)               # fine_tune.py:236
    import re
    pat = re.compile(r"model\.layers\.(\d+)\.")
    wg, warm_losses = {}, []
    for it in range(GI.CONFIG1["warmup_steps"]):
        model.zero_grad()
        out = model(input_ids=batches[it], labels=batches[it], use_cache=False)
        out.loss.backward()
        warm_losses.append(out.loss.item())
        for name, p in model.named_parameters():                 # fine_tune.py:716-768
            m = pat.search(name)
            ln = int(m.group(1)) if m else None
            if "self_attn" in name and "weight" in name:
                mn = "q_proj" if "q_proj" in name else "k_proj" if "k_proj" in name else "v_proj" if "v_proj" in name else None
                if mn is not None:
                    g = p.grad.detach().cpu().to(torch.float32)
                    wg[(mn, ln)] = g if (mn, ln) not in wg else wg[(mn, ln)] + g
    model.zero_grad()
    sel = H.select_submatrix_based_on_grads(wg, dims, n_attn, selection_strategy="no_restriction")
    model = S.freeze_unselected_matrix_layer(model, {}, sel)
    model = S.convert_linear_layer_to_matrix_sparsity(model, {}, sel)
    groups = S.get_optimizer_sparse_grouped_parameters(model, 0.0, GI.CONFIG1["smt_lr"])
    opt = torch.optim.AdamW(groups, lr=GI.CONFIG1["smt_lr"], betas=(0.9, 0.95), eps=1e-8, weight_decay=0.0)
    params = [p for g in groups for p in g["params"]]
    losses, grad_norms = [], []
    first_grads = None
    for it in range(GI.CONFIG1["sparse_steps"]):
        opt.zero_grad()
        b = batches[GI.CONFIG1["warmup_steps"] + it]
        out = model(input_ids=b, labels=b, use_cache=False)
        out.loss.backward()
        if first_grads is None:
            first_grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.requires_grad}
        grad_norms.append(float(torch.nn.utils.clip_grad_norm_(params, 1.0)))
        opt.step()
        losses.append(out.loss.item())
    final = {n: p.detach().clone() for n, p in model.named_parameters() if p.requires_grad}
    torch.save({"config": GI.CONFIG1, "dims": dims, "total_blocks": total, "n_attn": n_attn,
                "warm_losses": warm_losses, "selection": [(k, list(v)) for k, v in sel.items()],
                "trainable": [(n, tuple(p.shape)) for n, p in model.named_parameters() if p.requires_grad],
                "losses": losses, "grad_norms": grad_norms, "first_grads": first_grads,
                "final_selected": final}, os.path.join(OUT, "config1_e2e.pt"))
    print("config1:", dict(sel), "losses", losses, "norms", grad_norms)


def gen_config1_bf16(S):
    """Same model / batches as config 1 but in bf16 (the dtype the reference trains in, fine_tune.py --dtype bf16):
    the SMT steady state only — selection is TAKEN from the fp32 golden (bf16 noise could flip near-ties between
    two implementations, which would make a cross-implementation comparison of everything downstream meaningless).
    Stores losses, first-step gradients and the clipped-AdamW-updated compact weights after 4 steps.  The optimizer
    runs on fp32 masters of the bf16 compact parameters, as DeepSpeed's bf16 engine does (fine_tune.py:379-384)."""
    gold = torch.load(os.path.join(OUT, "config1_e2e.pt"), weights_only=False)
    sel = {k: [tuple(t) for t in v] for k, v in gold["selection"]}
    model, batches = GI.make_config1()
    model = model.to(torch.bfloat16)
    model = S.freeze_unselected_matrix_layer(model, {}, sel)
    model = S.convert_linear_layer_to_matrix_sparsity(model, {}, sel)
    named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
    masters = [p.detach().float().clone().requires_grad_(True) for _n, p in named]
    opt = torch.optim.AdamW(masters, lr=GI.CONFIG1["smt_lr"], betas=(0.9, 0.95), eps=1e-8, weight_decay=0.0)
    losses, first_grads = [], None
    for it in range(GI.CONFIG1["sparse_steps"]):
        for _n, p in named:
            p.grad = None
        b = batches[GI.CONFIG1["warmup_steps"] + it]
        out = model(input_ids=b, labels=b, use_cache=False)
        out.loss.backward()
        if first_grads is None:
            first_grads = {n: p.grad.detach().clone() for n, p in named}
        for m, (_n, p) in zip(masters, named):
            m.grad = p.grad.detach().float()
        torch.nn.utils.clip_grad_norm_(masters, 1.0)
        opt.step()
        with torch.no_grad():
            for m, (_n, p) in zip(masters, named):
                p.copy_(m.to(torch.bfloat16))
        losses.append(out.loss.item())
    torch.save({"selection": gold["selection"], "losses": losses, "first_grads": first_grads,
                "final_selected": {n: p.detach().clone() for n, p in named}},
               os.path.join(OUT, "config1_bf16.pt"))
    print("config1 bf16 losses", losses)


def main():
    os.makedirs(OUT, exist_ok=True)
    S, H = load_reference()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    gen_selection(H)
    gen_channels(H)
    gen_linearz(S)
    gen_linearchannel(S)
    gen_config1(S, H)
    gen_config1_bf16(S)


if __name__ == "__main__":
    main()
