"""Mirror of the reference's `smt` package (deepspeed/smt/): `smt.smt` and `smt.smt_helper`.

Put `<repo>` and `<repo>/sparse_matrix_tuning_b200` on PYTHONPATH and the reference driver's
`from smt.smt import ...` / `from smt.smt_helper import ...` (fine_tune.py:39-40) resolve to this package.

When it is reached under that top-level name, the package aliases itself to its canonical name
(`sparse_matrix_tuning_b200.smt`) so that there is exactly ONE instance of every module — the classes the driver
converts the model with are the classes `SMTAdam`, `dp` and `checkpoint` test against.
"""
import sys as _sys

if __name__ == "smt":
    import importlib as _importlib

    _canon = _importlib.import_module("sparse_matrix_tuning_b200.smt")
    _smt = _importlib.import_module("sparse_matrix_tuning_b200.smt.smt")
    _helper = _importlib.import_module("sparse_matrix_tuning_b200.smt.smt_helper")
    _sys.modules["smt"] = _canon
    _sys.modules["smt.smt"] = _smt
    _sys.modules["smt.smt_helper"] = _helper
