"""CPU: the C-ABI library loads, exports exactly what include/bgmm_b200.h declares, and fails loudly without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "bgmm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bgmm_[A-Za-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    from pybgmm_b200 import _lib
    names = _declared()
    assert len(names) >= 24
    L = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    assert sorted(_lib.EXPORTS) == names


def test_version_and_error_strings():
    from pybgmm_b200 import _lib
    assert b"sm_100a" in _lib.lib().bgmm_version()
    assert isinstance(_lib.lib().bgmm_last_error(), bytes)


def test_no_device_is_a_loud_error():
    """No CPU fallback: creating a chain without a CUDA device must raise (BGMM_ENODEV), never compute."""
    from pybgmm_b200 import _lib
    if _lib.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.BgmmError) as ei:
        _lib.Chain(np.zeros((4, 2)), np.zeros(2), 1.0, 3, np.eye(2), 4)
    assert ei.value.code == _lib.BGMM_ENODEV
    import pybgmm_b200 as P
    with pytest.raises(_lib.BgmmError):
        P.CRPMM(np.random.randn(8, 2), P.NIW(np.zeros(2), 1.0, 4, np.eye(2)), 1.0, None, K=2)


def test_argument_validation_before_device():
    from pybgmm_b200 import _lib
    L = _lib.lib()
    h = ctypes.c_void_p()
    X = np.zeros((4, 2))
    dp = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))  # noqa: E731
    m0, S0 = np.zeros(2), np.eye(2)
    rc = L.bgmm_create(dp(X), 4, 2, 0, dp(m0), 1.0, 1, dp(S0), 4, None, None, 0, 0, ctypes.byref(h))  # v0 < D
    assert rc == _lib.BGMM_EINVAL and b"v_0" in L.bgmm_last_error()
    rc = L.bgmm_create(dp(X), 4, 2, 7, dp(m0), 1.0, 3, dp(S0), 4, None, None, 0, 0, ctypes.byref(h))
    assert rc == _lib.BGMM_EINVAL
    rc = L.bgmm_create(dp(X), 4, 65, 0, dp(m0), 1.0, 70, dp(S0), 4, None, None, 0, 0, ctypes.byref(h))
    assert rc == _lib.BGMM_EINVAL
    assert L.bgmm_sweep(None, None, None, 1.0, 1.0, None) == _lib.BGMM_EINVAL


def test_oracle_is_not_imported_by_the_product():
    """The product path must not route through oracle/ (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "pybgmm_b200")
    for dirpath, _, files in os.walk(pkg):
        if "build" in dirpath.split(os.sep):
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), os.path.join(dirpath, f)
                assert "liboracle" not in text and "orc_" not in text, os.path.join(dirpath, f)


def test_tma_reductions_are_fp64_adds():
    """The statistics deltas go to global memory as TMA bulk reductions (cp.reduce.async.bulk ... .add.f64).  ptxas 12.9
    was seen to encode that PTX as an INTEGER add (UBLKRED.G.S.ADD.U64) in one of two kernels that shared the device
    function; the shipped SASS must hold the fp64 form only."""
    import shutil
    import subprocess
    from pybgmm_b200 import _lib
    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([tool, "-sass", _lib.LIB_PATH], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    kinds = set(re.findall(r"UBLKRED\S*", out))
    assert kinds, "no TMA reduction in the library"
    assert kinds == {"UBLKRED.G.S.ADD.F64.RN"}, kinds
