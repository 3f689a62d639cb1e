"""CPU: the C-ABI library loads, exports every symbol include/b200als.h declares, and refuses to
compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from rsparse_b200 import _lib as L

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "b200als.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200als_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = L.lib()
    names = _declared()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libb200als.so does not export %s" % n
    assert set(names) == set(L.exported_symbols())


def test_version_and_options_defaults():
    assert L.lib().b200als_version() == 100
    o = L.Options()
    L.lib().b200als_default_options(C.byref(o))
    assert (o.feedback, o.solver, o.cg_steps, o.dynamic_lambda) == (L.IMPLICIT, L.CONJUGATE_GRADIENT, 3, 1)


def test_library_does_not_link_the_oracle():
    import subprocess
    out = subprocess.run(["ldd", L.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out and "libref_wrmf" not in out


def test_compute_fails_loudly_without_gpu():
    if L.device_count() > 0:
        pytest.skip("a GPU is visible")
    from rsparse_b200 import als_implicit
    ptr = np.array([0, 1], np.int32)
    idx = np.array([0], np.int32)
    val = np.array([2.0])
    X = np.ones((1, 4), np.float32)
    Y = np.ones((1, 4), np.float32)
    with pytest.raises(L.B200AlsError) as e:
        als_implicit(ptr, idx, val, X, Y, 0.1, L.CONJUGATE_GRADIENT)
    assert e.value.code == L.ECUDA
    assert np.all(Y == 1)  # untouched: nothing was computed on the CPU


def test_synth_csr_host_generator():
    n_rows, n_cols, nnz = 1000, 5000, 37
    ptr = np.zeros(n_rows + 1, np.int32)
    idx = np.zeros(n_rows * nnz, np.int32)
    v32 = np.zeros(n_rows * nnz, np.float32)
    v64 = np.zeros(n_rows * nnz, np.float64)
    L.check(L.lib().b200als_synth_csr_host(n_rows, n_cols, nnz, 42, 0, 0, L.vp(ptr), L.vp(idx), L.vp(v32), L.vp(v64)))
    assert np.array_equal(ptr, np.arange(n_rows + 1) * nnz)
    rows = idx.reshape(n_rows, nnz)
    assert np.all(np.diff(rows, axis=1) > 0)          # ascending and distinct within a row
    assert rows.min() >= 0 and rows.max() < n_cols
    assert np.array_equal(v32.astype(np.float64), v64) and v32.min() >= 1 and v32.max() <= 10
    # a row-offset block reproduces the corresponding rows of the full matrix (sharding invariant)
    idx2 = np.zeros(100 * nnz, np.int32)
    ptr2 = np.zeros(101, np.int32)
    L.check(L.lib().b200als_synth_csr_host(100, n_cols, nnz, 42, 0, 300, L.vp(ptr2), L.vp(idx2), None, None))
    assert np.array_equal(idx2.reshape(100, nnz), rows[300:400])
    # explicit ratings
    L.check(L.lib().b200als_synth_csr_host(n_rows, n_cols, nnz, 42, 1, 0, L.vp(ptr), L.vp(idx), L.vp(v32), None))
    assert set(np.unique(v32)) <= {1.0, 2.0, 3.0, 4.0, 5.0}


def test_oracle_generator_equals_the_products():
    """bench.py's CPU arms build their CSR with oracle.synth_csr so that the reference arm never loads the product
    library; both generators must yield the same matrix, entry for entry (implicit, explicit, row offsets)."""
    import oracle
    n_rows, n_cols, nnz = 777, 4001, 23
    for explicit, off in ((0, 0), (1, 0), (0, 12345)):
        ptr = np.zeros(n_rows + 1, np.int32)
        idx = np.zeros(n_rows * nnz, np.int32)
        v64 = np.zeros(n_rows * nnz, np.float64)
        L.check(L.lib().b200als_synth_csr_host(n_rows, n_cols, nnz, 42, explicit, off, L.vp(ptr), L.vp(idx), None, L.vp(v64)))
        p2, i2, v2 = oracle.synth_csr(n_rows, n_cols, nnz, 42, bool(explicit), off)
        assert np.array_equal(ptr, p2) and np.array_equal(idx, i2) and np.array_equal(v64, v2)


def test_bad_arguments_are_rejected():
    assert L.lib().b200als_synth_csr_host(10, 5, 6, 1, 0, 0, None, None, None, None) == L.EINVAL
    assert b"synthetic" in L.lib().b200als_last_error()


def test_r_shim_type_checks_against_the_abi():
    """r-package/src/b200als_shim.c (the `.Call` layer for an R build) is type-checked against include/b200als.h with
    stand-in R headers carrying R's real signatures: every ABI call in the shim matches the header, -Werror."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    p = subprocess.run(["sh", os.path.join(ROOT, "r-package", "tests", "check_shim.sh")], capture_output=True, text=True)
    assert p.returncode == 0 and "type-check ok" in p.stdout, p.stdout + p.stderr
