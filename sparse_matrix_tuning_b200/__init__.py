"""B200-native implementation of the SMT (Sparse Matrix Tuning) hot path.

    sparse_matrix_tuning_b200.smt.smt / .smt.smt_helper   drop-in mirror of the reference's `smt` package
    sparse_matrix_tuning_b200.ops                         tensor-level wrappers over the C-ABI (include/smt_b200.h)
    sparse_matrix_tuning_b200.optim.SMTAdam               fused compact AdamW + clip + dense write-back
    sparse_matrix_tuning_b200.warmup                      on-device warm-up score accumulation
    sparse_matrix_tuning_b200.dp                          data-parallel exchange of the compact buffers (NCCL)

All arithmetic runs in `_C/libsmt_b200.so` (hand-written sm_100a CUDA). There is no CPU fallback.
"""
__version__ = "0.1.0"
