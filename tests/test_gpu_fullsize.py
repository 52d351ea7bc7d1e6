"""GPU parity at the FULL BASELINE.json sizes (configs 3, 4, 5a, 5b) against the oracle itself.

BaseTCSC is row-independent (comp.h:37-63: the m loop is outermost and nothing crosses rows), so the
oracle on a handful of rows of X is the oracle's answer for those rows of the full call.  Per shape:
  * W is drawn on the GPU (K·N up to 4.7e8), copied to the host once, and the ORACLE builds its own
    TCSC from it (TCSC.h:13-41): the device-built arrays must equal them bit for bit — full-size
    format parity, not a round trip;
  * the full-M call runs on the GPU through the C ABI with the kernel AUTO picks for that M, integer
    X (initX regime: bit-exact) and real-valued X (U(-1,1): the 1e-5 bar of BASELINE.json, applied
    as max-norm relative error AND as the element-wise error against the rigorous forward bound
    (n+2)·eps·(Σ|x|+|b|); both figures are printed);
  * 8 random rows are compared with orc.base_tcsc / orc.base_tcsc_prelu; the other kernels AUTO can
    fall to (gather, code_gemv for M <= 2) are run on those rows as well.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5
EPS = np.finfo(np.float32).eps

CASES = [  # name, M, K, N, s, prelu
    ("c3", 256, 4096, 14336, 4, True),
    ("c4", 2048, 8192, 28672, 8, False),
    ("c5a", 32, 8192, 57344, 4, False),
    ("c5b", 512, 8192, 57344, 4, False),
]


def _bound(X, o, b):
    """(n+2)·eps·(Σ_{k in col} |x_k| + |b|) per element, from the oracle's own TCSC."""
    csp, csn, rip, rin = o.arrays
    absX = np.abs(X).astype(np.float64)
    cum = np.concatenate([np.zeros((X.shape[0], 1)), np.cumsum(absX[:, rip], axis=1)], axis=1)
    tot = cum[:, csp[1:]] - cum[:, csp[:-1]]
    cum = np.concatenate([np.zeros((X.shape[0], 1)), np.cumsum(absX[:, rin], axis=1)], axis=1)
    tot += cum[:, csn[1:]] - cum[:, csn[:-1]]
    cnt = (np.diff(csp) + np.diff(csn)).astype(np.float64)
    return (cnt[None, :] + 2) * EPS * (tot + np.abs(b)[None, :].astype(np.float64))


@pytest.mark.parametrize("name,M,K,N,s,prelu", CASES, ids=[c[0] for c in CASES])
def test_full_size_against_oracle(tsg, orc, name, M, K, N, s, prelu, capsys):
    import torch
    from ternary_spgemm_b200 import synth
    Wd = synth.device_ternary(K, N, s, 2024)
    t = tsg.TCSC.from_device_dense(Wd, K, N, elem_bytes=1)
    W = Wd.cpu().numpy().astype(np.int32)
    del Wd
    torch.cuda.empty_cache()
    o = orc.tcsc(W)                                        # the oracle's own constructor, full size
    del W
    for got, exp, nm in zip(t.export(), o.arrays, ("csp", "csn", "rip", "rin")):
        assert got.shape == exp.shape and np.array_equal(got, exp), f"{name}: {nm} differs from the oracle"
    rng = np.random.default_rng(7)
    rows = np.sort(rng.choice(M, size=min(8, M), replace=False))
    b = rng.uniform(-2, 2, N).astype(np.float32)
    al = rng.uniform(0.01, 0.3, N).astype(np.float32) if prelu else None
    auto = tsg.ALGO_NAMES[t.pick(M)]
    report = []
    for regime in ("int", "real", "real_fast"):
        # real_fast: the opt-in two-fp16-term tiles (tsg_set_fast_split, include/tsg.h) on the same X
        tsg.set_fast_split(regime == "real_fast")
        if regime != "real_fast":
            X = (rng.integers(-512, 513, (M, K)).astype(np.float32) if regime == "int"
                 else rng.uniform(-1, 1, (M, K)).astype(np.float32))
        Y = t.spmm(X, b, al)                               # full M, the kernel AUTO picks
        Xs = np.ascontiguousarray(X[rows])
        want = orc.base_tcsc_prelu(Xs, o, b, al) if prelu else orc.base_tcsc(Xs, o, b)
        others = {"gather": t.spmm(Xs, b, al, algo=tsg.ALGO_GATHER),
                  "dense_tc@8rows": t.spmm(Xs, b, al, algo=tsg.ALGO_DENSE_TC),
                  "code_gemv@2rows": t.spmm(Xs[:2], b, al, algo=tsg.ALGO_CODE_GEMV)}
        checks = {f"auto({auto})@M={M}": (Y[rows], want)}
        for k, v in others.items():
            checks[k] = (v, want[: v.shape[0]])
        if regime == "int":
            for k, (got, exp) in checks.items():
                assert np.array_equal(got, exp), f"{name} {k}: integer X must be bit-identical to BaseTCSC"
            report.append(f"{name} int X: bit-identical ({', '.join(checks)})")
            continue
        bound = _bound(Xs, o, b)
        if prelu:                                          # |alpha| < 1: the bound carries through PReLU
            pass
        for k, (got, exp) in checks.items():
            err = np.abs(got.astype(np.float64) - exp.astype(np.float64))
            maxnorm = err.max() / max(float(np.abs(exp).max()), 1e-30)
            nz = np.abs(exp) > 0
            elem = float((err[nz] / np.abs(exp[nz])).max())
            med = float(np.median(err[nz] / np.abs(exp[nz])))
            frac_bound = float((err / bound[: got.shape[0]]).max())
            report.append(f"{name} {regime} X {k}: max-norm rel {maxnorm:.2e}, element-wise rel max {elem:.2e} "
                          f"(median {med:.2e}), max err / forward bound {frac_bound:.3f}")
            assert maxnorm <= REL_TOL, (name, k, maxnorm)
            assert frac_bound <= 1.0, (name, k, frac_bound)
    tsg.set_fast_split(False)
    with capsys.disabled():
        print("\n" + "\n".join(report))
