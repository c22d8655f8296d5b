"""Mirror of the reference's `smt` package (deepspeed/smt/): `smt.smt` and `smt.smt_helper`.

Put `<repo>/sparse_matrix_tuning_b200` (next to the repo root itself) on PYTHONPATH and the reference driver's
`from smt.smt import ...` / `from smt.smt_helper import ...` (fine_tune.py:39-40) resolve to this package.
"""
