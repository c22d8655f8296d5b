"""Import shim that lets the UNMODIFIED reference modules (`/root/reference/deepspeed/smt/{smt,smt_helper}.py`)
import in a container without deepspeed / pytorch_memlab / matplotlib — TEST INFRASTRUCTURE ONLY.

Used by `oracle/gen_golden.py` (to produce the committed fixtures under `tests/golden/`) and by the CPU
tests that pin `oracle/smt_oracle.py` against the live reference.  `/root/reference` does not exist on the
GPU box, so nothing on the GPU path may call `load_reference()`.

What is faked (nothing of the SMT algorithm itself):
  deepspeed.init_distributed            -> single-process gloo process group (smt.py:20 runs it at import)
  deepspeed.compression.helper          -> recursive_getattr / recursive_setattr (dotted-name walk)
  deepspeed.ops.adam                    -> FusedAdam / DeepSpeedCPUAdam = torch.optim.AdamW
  deepspeed.ops.adam.multi_tensor_apply, deepspeed.runtime.utils, deepspeed.accelerator, deepspeed.comm,
  pytorch_memlab, matplotlib(.pyplot), huggingface_hub.snapshot_download  -> inert placeholders
"""
from __future__ import annotations

import os
import sys
import types

# `oracle/_ref/` = the two reference modules staged, unmodified, by `oracle/build_ref.py` (git-ignored, but it travels to
# the GPU box with the snapshot): lets `bench.py` time the reference ITSELF where `/root/reference` does not exist.
STAGED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def _pick_root() -> str:
    env = os.environ.get("SMT_REFERENCE_ROOT")
    for cand in (env, STAGED_ROOT, "/root/reference"):
        if cand and os.path.isfile(os.path.join(cand, "deepspeed", "smt", "smt.py")):
            return cand
    return env or "/root/reference"


REFERENCE_ROOT = _pick_root()


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "deepspeed", "smt", "smt.py"))


def _mk(name: str) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__["__smt_shim__"] = True
    sys.modules[name] = m
    return m


def _install_fakes() -> None:
    import torch

    if "deepspeed" in sys.modules and not getattr(sys.modules["deepspeed"], "__smt_shim__", False):
        return  # a real deepspeed is installed: use it
    ds = _mk("deepspeed")

    def init_distributed(*_a, **_k):
        if not torch.distributed.is_initialized():
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            os.environ.setdefault("MASTER_PORT", str(29500 + os.getpid() % 2000))
            torch.distributed.init_process_group("gloo", rank=0, world_size=1)

    ds.init_distributed = init_distributed
    _mk("deepspeed.compression")
    helper = _mk("deepspeed.compression.helper")

    def recursive_getattr(model, name):
        out = model
        for part in name.split("."):
            out = getattr(out, part)
        return out

    def recursive_setattr(model, name, module):
        parts = name.split(".")
        out = model
        for part in parts[:-1]:
            out = getattr(out, part)
        setattr(out, parts[-1], module)

    helper.recursive_getattr = recursive_getattr
    helper.recursive_setattr = recursive_setattr
    _mk("deepspeed.ops")
    adam = _mk("deepspeed.ops.adam")
    adam.FusedAdam = torch.optim.AdamW
    adam.DeepSpeedCPUAdam = torch.optim.AdamW
    mta = _mk("deepspeed.ops.adam.multi_tensor_apply")
    mta.MultiTensorApply = object
    _mk("deepspeed.runtime")
    ru = _mk("deepspeed.runtime.utils")
    ru.see_memory_usage = lambda *a, **k: None
    acc = _mk("deepspeed.accelerator")
    acc.get_accelerator = lambda: None
    comm = _mk("deepspeed.comm")
    ds.comm = comm
    _mk("deepspeed.runtime.zero")
    zpp = _mk("deepspeed.runtime.zero.partition_parameters")
    zpp.ZeroParamStatus = object
    ml = _mk("pytorch_memlab")
    ml.MemReporter = object
    try:
        import matplotlib.pyplot  # noqa: F401
    except Exception:
        mp = _mk("matplotlib")
        mp.pyplot = _mk("matplotlib.pyplot")


_cached = None


def load_reference():
    """Returns (smt_module, smt_helper_module) of the reference, imported under private names so they cannot
    shadow (or be shadowed by) this repo's own `smt` mirror package."""
    global _cached
    if _cached is not None:
        return _cached
    if not reference_available():
        raise FileNotFoundError(f"reference not found under {REFERENCE_ROOT}")
    import importlib.util

    _install_fakes()
    sys.dont_write_bytecode = True
    ds_dir = os.path.join(REFERENCE_ROOT, "deepspeed")

    # `from helpers.deepspeed_helpers import print_rank_0` (smt.py:17): provide just that symbol, the real
    # helpers module drags in huggingface_hub / HfDeepSpeedConfig and is out of scope.
    helpers = _mk("helpers")
    dh = _mk("helpers.deepspeed_helpers")

    def print_rank_0(msg, rank=None):
        if rank is not None and rank <= 0:
            print(msg)
        elif rank is None:
            print(msg)

    dh.print_rank_0 = print_rank_0
    helpers.deepspeed_helpers = dh

    mods = []
    for fname, alias in (("smt.py", "_ref_smt_smt"), ("smt_helper.py", "_ref_smt_helper")):
        spec = importlib.util.spec_from_file_location(alias, os.path.join(ds_dir, "smt", fname))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[alias] = mod
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            spec.loader.exec_module(mod)
        mods.append(mod)
    # remove the placeholder `helpers` package again so it cannot leak into other imports
    for k in ("helpers", "helpers.deepspeed_helpers"):
        sys.modules.pop(k, None)
    _cached = tuple(mods)
    return _cached
