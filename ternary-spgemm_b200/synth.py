"""Synthetic workloads at the BASELINE.json shapes (torch is plumbing here: device memory + RNG).

`device_ternary` draws W on the GPU with the same row statistics as the reference generator
generateSparseMatrix<int>(K, N, s, false) (cpp_impl/sparseUtils.h:52-87): every row holds
2*((N//s)//2) non-zeros at distinct uniformly random columns, (N//s)//2 + v of them +1 and the
rest -1 with v ~ U{0 .. N//s//20 + 1}.  It is NOT bit-identical to the mt19937 stream (parity
tests use the real generator through the checker); it exists so that K·N up to 4.7e8 can be
generated in HBM in milliseconds instead of seconds on one host core.
"""
from __future__ import annotations

CONFIGS = {
    # BASELINE.json configs[0..4]
    "c1": dict(M=32, K=1024, N=4096, s=4, prelu=False, note="README example (CPU-runnable)"),
    "c2": dict(M=1, K=4096, N=4096, s=3, prelu=False, note="GEMV-style decode shape"),
    "c3": dict(M=256, K=4096, N=14336, s=4, prelu=True, note="BitNet-style FFN up-proj, bias+PReLU"),
    "c4": dict(M=2048, K=8192, N=28672, s=8, prelu=False, fmt="pcsc", note="packed-value CSC, N-sharded 2/4/8 GPUs"),
    "c5a": dict(M=32, K=8192, N=57344, s=4, prelu=False, note="sparsity sweep point, M=32"),
    "c5b": dict(M=512, K=8192, N=57344, s=4, prelu=False, note="sparsity sweep point, M=512"),
}


def flops(M: int, N: int, K: int, s: int) -> float:
    """The reference's flop model M·N·(1 + K/s) (readme.md:84-85)."""
    return M * N * (1.0 + K / s)


def device_ternary(K: int, N: int, s: int, seed: int, device="cuda"):
    """int8 K×N ternary matrix in HBM (row-major)."""
    import torch

    g = torch.Generator(device=device).manual_seed(seed)
    W = torch.zeros(K, N, dtype=torch.int8, device=device)
    per = N // s
    half = per // 2
    cnt = 2 * half
    if cnt == 0:
        return W
    vmax = per // 20 + 1
    rows_per = max(1, (1 << 24) // N)
    ar = torch.arange(cnt, device=device)
    for k0 in range(0, K, rows_per):
        k1 = min(K, k0 + rows_per)
        cols = torch.rand(k1 - k0, N, device=device, generator=g).argsort(dim=1)[:, :cnt]
        v = torch.randint(0, vmax + 1, (k1 - k0, 1), device=device, generator=g)
        npos = torch.clamp(half + v, max=cnt)
        sign = torch.where(ar[None, :] < npos, 1, -1).to(torch.int8)
        W[k0:k1].scatter_(1, cols, sign)
    return W


def device_x(M: int, K: int, seed: int, device="cuda", integer: bool = True):
    """X in the reference's regime (integers in [-512, 512] stored as fp32, sparseUtils.h:6-23)."""
    import torch

    g = torch.Generator(device=device).manual_seed(seed)
    if integer:
        return torch.randint(-512, 513, (M, K), device=device, generator=g).to(torch.float32)
    return torch.rand(M, K, device=device, generator=g) * 2 - 1
