"""CPU: libisb.so loads, exports every symbol include/isb.h declares, and its
argument checks / error reporting work without a GPU (no compute calls)."""

import ctypes
import os
import re

import pytest

from conftest import ROOT


@pytest.fixture(scope="module")
def L():
    from instance_search_b200 import _lib
    return _lib.lib()


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "isb.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(isb_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(L):
    from instance_search_b200 import _lib
    names = _declared_symbols()
    assert len(names) >= 18
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), "libisb.so does not export %s" % n
        assert n in _lib.SIGNATURES, "ctypes binding of %s missing in _lib.SIGNATURES" % n
    # and nothing is bound that the header does not declare
    assert sorted(_lib.SIGNATURES) == names


def test_abi_version(L):
    assert L.isb_abi_version() == 6


def test_argument_errors_need_no_gpu(L):
    # status codes + thread-local message (include/isb.h conventions)
    rc = L.isb_l2norm_rows(None, -1, 4, 1e-10, None, None)
    assert rc == 1 and b"negative" in L.isb_last_error()
    rc = L.isb_topk_merge(None, None, 1, 1, 1, None, None, None)
    assert rc == 1 and b"null pointer" in L.isb_last_error()
    rc = L.isb_f32_to_bf16(ctypes.c_void_p(16), 1, 8, 8, ctypes.c_void_p(16), 12, 0, None)
    assert rc == 1 and b"multiple of 8" in L.isb_last_error()
    rc = L.isb_gemm_nt(ctypes.c_void_p(16), 8, ctypes.c_void_p(16), 8, 0, 1, 8, None, ctypes.c_void_p(16),
                       1, 1, None, 0, None)
    assert rc == 1 and b"empty" in L.isb_last_error()
    # k + margin beyond the candidate budget
    rc = L.isb_topk_search(ctypes.c_void_p(16), 1, ctypes.c_void_p(16), ctypes.c_void_p(16), 1000, 64, 64,
                           100, 29, 0, ctypes.c_void_p(16), ctypes.c_void_p(16), None, None, None, 0, None)
    assert rc == 1 and b"k + margin" in L.isb_last_error()


def test_no_device_is_an_error_not_a_fallback(L):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert L.isb_check_device() == 4          # ISB_ERR_UNSUPPORTED_DEVICE
    assert L.isb_last_error()
    # a well-formed search call fails loudly on the device check
    rc = L.isb_topk_search(ctypes.c_void_p(16), 1, ctypes.c_void_p(16), ctypes.c_void_p(16), 1000, 64, 64,
                           10, 5, 0, ctypes.c_void_p(16), ctypes.c_void_p(16), None, None,
                           ctypes.c_void_p(1024), 1 << 30, None)
    assert rc == 4


def test_workspace_queries(L):
    n = L.isb_topk_search_workspace_bytes(10000, 1000000, 2048, 100, 28)
    assert 100e6 < n < 2e9
    assert L.isb_topk_search_workspace_bytes(0, 10, 8, 1, 0) == 0
    assert L.isb_gemm_nt_workspace_bytes(256, 2048, 100352, 1) == 0
    assert L.isb_gemm_nt_workspace_bytes(256, 2048, 100352, 4) == 4 * 256 * 2048 * 4
    assert L.isb_region_select_workspace_bytes(2, 2048, 14, 14, 464, 7, 7, 6, 10) > 2 * 64 * 2048 * 2
    assert L.isb_select_negatives_workspace_bytes(100, 1000, 128, 1) > L.isb_select_negatives_workspace_bytes(100, 1000, 128, 0) > 0
    assert L.isb_topk_resolve_workspace_bytes(5, 100000, 256) > 0
    assert L.isb_topk_exhaustive_workspace_bytes(5, 100000, 100) >= 5 * 32 * 100 * 16


def test_ops_reject_cpu_tensors():
    import torch
    from instance_search_b200 import IsbError, ops
    with pytest.raises(IsbError):
        ops.l2norm_rows(torch.zeros(2, 8))
    with pytest.raises(IsbError):
        ops.topk_merge(torch.zeros(2, 3, 4), torch.zeros(2, 3, 4, dtype=torch.int64))


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "instance_search_b200")
    for dp, _, files in os.walk(pkg):
        for fn in files:
            if not fn.endswith(".py"):
                continue
            with open(os.path.join(dp, fn)) as f:
                src = f.read()
            m = re.search(r"^\s*(?:import|from)\s+oracle\b.*$", src, flags=re.M)
            assert m is None, "%s imports the oracle: %s" % (fn, m.group(0))


def test_shipped_library_is_not_the_timeline_debug_build():
    # make TIMELINE=1 (role timeline of the CTA-pair screen) is a debug build: the library the
    # tests, the bench and the driver load must be the plain one
    from instance_search_b200 import _lib
    with pytest.raises(AttributeError):
        _lib.lib().isb_debug_timeline


def test_options_are_explicit_not_environment(L, monkeypatch):
    # the library reads nothing from the environment (round-1 A/B switches were getenv calls)
    from instance_search_b200 import _lib
    monkeypatch.setenv("ISB_MINING_KC", "8")
    assert _lib.get_option("mining_kc") is None
    with _lib.options(mining_kc=8, screen_pair=0):
        assert _lib.get_option("mining_kc") == 8 and _lib.get_option("screen_pair") == 0
    assert _lib.get_option("mining_kc") is None and _lib.get_option("screen_pair") is None
    assert L.isb_set_option(999, 1) == 1 and b"unknown option" in L.isb_last_error()
    with pytest.raises(_lib.IsbError):
        _lib.set_option("no_such_option", 1)
    # (the static CUDA runtime inside the .so imports getenv, so the check is on our sources)
    import glob
    import os
    csrc = os.path.join(os.path.dirname(_lib.LIB_PATH), "csrc")
    for f in glob.glob(os.path.join(csrc, "*.cu")) + glob.glob(os.path.join(csrc, "*.cuh")):
        assert "getenv" not in open(f).read(), f
