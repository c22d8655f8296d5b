"""Stages the reference's own implementation of the hot path under `oracle/_ref/` - TEST / BENCH INFRASTRUCTURE ONLY.

The reference is pure Python (no native code, nothing to compile): "building" it means copying the two UNMODIFIED
modules that hold the path, `deepspeed/smt/smt.py` and `deepspeed/smt/smt_helper.py`, from where they lie under
`/root/reference` into the git-ignored directory `oracle/_ref/deepspeed/smt/`.  That directory is NOT part of the
repository's history (see .gitignore) but it is part of the `gpurun` snapshot, so the GPU box - which has no
`/root/reference` - can import the reference itself through `oracle/ref_shim.py`:

  * `bench.py --impl reference` and the `cpu_baseline` leg time the reference's `LinearLayer_MatrixSparsity` /
    `linearZ` on the host cores (`cpu_baseline.kind = "reference"`);
  * `bench.py`'s `secondary_comparator` runs the same reference modules eagerly on the B200.

Run by `__graft_entry__.build()` whenever `/root/reference` exists; a no-op otherwise (the prebuilt copy is used).
"""
from __future__ import annotations

import hashlib
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_ROOT = "/root/reference"
DST_ROOT = os.path.join(HERE, "_ref")
FILES = ("deepspeed/smt/smt.py", "deepspeed/smt/smt_helper.py")


def build_ref(verbose: bool = True) -> bool:
    """Returns True when `oracle/_ref` holds the reference modules afterwards."""
    if os.path.isfile(os.path.join(SRC_ROOT, FILES[0])):
        digests = []
        for rel in FILES:
            dst = os.path.join(DST_ROOT, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(os.path.join(SRC_ROOT, rel), dst)
            digests.append(f"{rel} sha256={hashlib.sha256(open(dst, 'rb').read()).hexdigest()[:16]}")
        with open(os.path.join(DST_ROOT, "STAGED_FROM.txt"), "w") as f:
            f.write("unmodified copies staged by oracle/build_ref.py from /root/reference\n" + "\n".join(digests) + "\n")
        if verbose:
            print("[build_ref] staged " + ", ".join(digests))
    return os.path.isfile(os.path.join(DST_ROOT, FILES[0]))


if __name__ == "__main__":
    print(build_ref())
