"""CPU: the C-ABI library loads without a GPU and exports every symbol include/tsg.h declares;
compute entry points fail loudly (no CPU fallback) when no device is present."""
import ctypes
import os
import re

import numpy as np
import pytest


def declared_symbols(header_path):
    src = open(header_path).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tsg_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(tsg):
    syms = declared_symbols(tsg.HEADER_PATH)
    assert len(syms) >= 20, syms
    L = ctypes.CDLL(tsg.LIB_PATH)
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, f"declared in tsg.h but not exported: {missing}"
    assert L.tsg_abi_version() == 1


def test_no_cpu_fallback(tsg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert tsg.device_count() == 0
    W = np.zeros((4, 4), np.int32)
    with pytest.raises(tsg.TsgError) as e:
        tsg.TCSC(W)
    assert e.value.status == -3  # TSG_ERR_NO_DEVICE


def test_product_never_touches_oracle():
    """The product tree must not reference oracle/ in any way."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "ternary-spgemm_b200")
    offenders = []
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")) or f == "Makefile":
                txt = open(os.path.join(d, f), errors="replace").read()
                if re.search(r"oracle|liboracle|libtsgref|pyoracle", txt):
                    offenders.append(os.path.join(d, f))
    assert not offenders, offenders


def test_shard_columns(tsg):
    for N in (1, 7, 4096, 28672, 57344):
        for world in (1, 2, 3, 4, 8):
            cuts = [tsg.shard_columns(N, world, r) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == N
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_formats_fail_loudly_without_gpu(tsg):
    """TCSR / packed CSC / BlockedTCSC go through the same device checks: no CPU fallback."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    W = np.zeros((4, 4), np.int32)
    for cls in (tsg.TCSR, tsg.PackedCSC, tsg.PackedCSR):
        with pytest.raises(tsg.TsgError) as e:
            cls(W)
        assert e.value.status == -3  # TSG_ERR_NO_DEVICE


def test_algo_enum_matches_header(tsg):
    """The Python constants mirror enum tsg_algo in include/tsg.h."""
    src = open(tsg.HEADER_PATH).read()
    enum = dict((n, int(v)) for n, v in re.findall(r"TSG_ALGO_([A-Z_]+)\s*=\s*(\d+)", src))
    assert enum == {"AUTO": tsg.ALGO_AUTO, "GATHER": tsg.ALGO_GATHER, "GATHER_SEQ": tsg.ALGO_GATHER_SEQ,
                    "DENSE_TC": tsg.ALGO_DENSE_TC, "CODE_GEMV": tsg.ALGO_CODE_GEMV,
                    "TCSR_SEQ": tsg.ALGO_TCSR_SEQ, "PCSC_GATHER": tsg.ALGO_PCSC_GATHER,
                    "PCSR_SEQ": tsg.ALGO_PCSR_SEQ}
    assert set(tsg.ALGO_NAMES) == set(enum.values())


def test_header_is_plain_c(tmp_path):
    """include/tsg.h is the drop-in boundary: it must compile as C99 (no C++, no CUDA, no torch
    types in the signatures) and link against libtsg.so from a C translation unit."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "abi.c"
    src.write_text('#include "tsg.h"\n#include <stdio.h>\n'
                   'int main(void) { int n = -1; int s = tsg_device_count(&n);\n'
                   '  printf("%d %d %d\\n", tsg_abi_version(), s, n); return 0; }\n')
    exe = tmp_path / "abi.out"
    lib = os.path.join(root, "ternary-spgemm_b200")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", f"-I{root}/include", str(src),
                        "-o", str(exe), f"-L{lib}", "-ltsg", f"-Wl,-rpath,{lib}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0
    ver, status, count = (int(v) for v in out.stdout.split())
    assert ver == 1 and status == 0 and count >= 0          # no GPU here: zero usable devices, not an error


def test_tail_plan_host_logic(tsg):
    """The tail launch of the tensor-core path is planned on the host: the tiles of a
    partial last wave, K-split over min(8, SMs / tiles) CTAs, only where it pays.  No GPU needed."""
    import ctypes as C
    L = tsg.lib()
    L.tsg_debug_plan_tail.argtypes = [C.c_longlong, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
    L.tsg_debug_plan_tail.restype = C.c_int

    def plan(tiles, nst, sms=148, nt=256):
        ks = C.c_int(0)
        return L.tsg_debug_plan_tail(tiles, nst, sms, nt, C.byref(ks)), ks.value

    assert plan(1792, 32) == (16, 8)        # c4: 12 waves + 16 tiles, eight-way split
    assert plan(448, 32) == (4, 8)          # its 4-GPU shard: 3 waves + 4 tiles
    assert plan(896, 32) == (8, 8)          # c5b / the 2-GPU shard
    assert plan(148 * 5, 32) == (0, 1)      # whole waves: nothing to do
    assert plan(112, 16) == (0, 1)          # c3: more tiles than half the SMs cannot be K-split evenly
    assert plan(148 + 50, 32) == (50, 2)    # two SMs per tile
    assert plan(20, 2) == (0, 1)            # two stages: the saving does not pay for the trip through L2
    assert plan(20, 64)[1] == 7             # all tail: 148 / 20
    assert plan(3, 4)[1] in (0, 1, 4)       # never more ranks than stages
