"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libtsgref.so).

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
Every array below is an output of reference code (generateSparseMatrix, TCSC::TCSC,
BaseTCSC<float>, BaseTCSC_PreLU<float>, DoubleUnrolledTCSC<float,4,4>, GEMM) on the inputs
stored next to it, so the fixtures pin both the oracle restatement (CPU tests) and the CUDA
path (GPU tests) on boxes where /root/reference does not exist.

Shapes: the reference's own tiny verbose case (K=3,N=4,s=2,seed=0,
cpp_impl/test_data_structure.cpp:149), ragged/odd shapes, its smallest "required" shape
(K=512,N=2048, test_data_structure.cpp:116-118) at s=2 and 16, and a 1/8-height slice of the
README example (config 1).  X is stored in both regimes: the reference's integer-valued initX
regime and real-valued U(-1,1).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.pyoracle import Oracle, Reference  # noqa: E402

CASES = [  # name, M, K, N, s, seed
    ("tiny_3x4", 2, 3, 4, 2, 0),
    ("ragged_37x29", 3, 37, 29, 3, 5),
    ("odd_100x130", 5, 100, 130, 4, 7),
    ("req_512x2048_s2", 4, 512, 2048, 2, 0),
    ("req_512x2048_s16", 4, 512, 2048, 16, 1),
    ("c1_slice_128x4096_s4", 8, 128, 4096, 4, 1234),
]


def main():
    ref, orc = Reference(), Oracle()
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name, M, K, N, s, seed in CASES:
        W = ref.generate_sparse_matrix(K, N, s, seed)
        t = ref.tcsc(W)
        h = ref.tcsc_handle(W)
        rng = np.random.default_rng(1000 + seed)
        X_int = orc.init_x(M, K, 4242 + seed)          # reference regime: integers in [-512,512]
        X_real = rng.uniform(-1.0, 1.0, (M, K)).astype(np.float32)
        b = rng.uniform(-2.0, 2.0, N).astype(np.float32)
        alpha = rng.uniform(0.01, 0.3, N).astype(np.float32)
        b_ref = np.full(N, 2.0, np.float32)             # main.cpp:194-195
        a_ref = np.full(N, 0.1, np.float32)
        np.savez_compressed(
            os.path.join(out_dir, name + ".npz"),
            shape=np.array([M, K, N, s, seed], np.int64),
            W=W.astype(np.int8),
            csp=t.col_start_pos, csn=t.col_start_neg, rip=t.row_index_pos, rin=t.row_index_neg,
            ds_bytes=np.int64(ref.tcsc_size_bytes(h)),
            X_int=X_int, X_real=X_real, b=b, alpha=alpha,
            Y_int=ref.base_tcsc(h, X_int, b_ref),
            Y_int_prelu=ref.base_tcsc_prelu(h, X_int, b_ref, a_ref),
            Y_int_dense=ref.gemm(X_int, W, b_ref),
            Y_real=ref.base_tcsc(h, X_real, b),
            Y_real_prelu=ref.base_tcsc_prelu(h, X_real, b, alpha),
            Y_real_du44=ref.double_unrolled_tcsc_k4_m4(h, X_real, b),
        )
        print("wrote", name, "nnz", t.nnz)


if __name__ == "__main__":
    main()
