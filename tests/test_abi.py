"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol include/paos_b200.h
declares, the ctypes binding covers them all, and without a GPU every entry point fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    text = open(os.path.join(ROOT, "include", "paos_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(paos_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from paos_b200 import _lib

    names = declared_functions()
    assert len(names) >= 25
    for name in names:
        assert hasattr(_lib.lib, name), f"{name} declared in include/paos_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes binding and header disagree"
    assert _lib.lib.paos_abi_version() == _lib.ABI_VERSION


def test_no_cpu_fallback_without_device():
    import paos_b200
    from paos_b200 import _lib

    if paos_b200.device_count() > 0:
        pytest.skip("a B200 is present")
    h = C.c_void_p()
    rc = _lib.lib.paos_wfo_create(C.byref(h), 64, _lib.PAOS_C128, 0, None, None)
    assert rc == _lib.PAOS_ERR_CUDA and not h.value
    assert b"no CPU fallback" in _lib.lib.paos_last_error() or b"CUDA" in _lib.lib.paos_last_error()
    with pytest.raises(paos_b200.PaosCudaError):
        paos_b200.WFO(1.0, 1e-6, 64, 4)
    job_args = (1.0, 1e-6, 64, 4, {"us": 0.0, "ut": 0.0}, {})
    with pytest.raises(paos_b200.PaosCudaError):
        paos_b200.run(*job_args)


def test_argument_validation_is_reported_not_crashed():
    from paos_b200 import _lib

    h = C.c_void_p()
    assert _lib.lib.paos_wfo_create(C.byref(h), 100, _lib.PAOS_C128, 0, None, None) == _lib.PAOS_ERR_ARG
    assert _lib.lib.paos_wfo_create(C.byref(h), 64, 7, 0, None, None) == _lib.PAOS_ERR_ARG
    assert _lib.lib.paos_wfo_create(None, 64, 0, 0, None, None) == _lib.PAOS_ERR_ARG
    assert _lib.lib.paos_wfo_flush(None) == _lib.PAOS_ERR_ARG
    assert _lib.lib.paos_wfo_destroy(None) == _lib.PAOS_OK


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "paos_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f"{f} imports the oracle"
                assert "cufft" not in text.lower(), f"{f} mentions cuFFT"


def test_library_is_built_for_sm100a_only():
    import subprocess

    lib = os.path.join(ROOT, "paos_b200", "libpaos_b200.so")
    try:
        out = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
    except OSError:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_ctypes_mirrors_have_the_compiled_layout():
    """paos_surface / paos_snapshot / paos_stats as mirrored in Python have the size the library was compiled with."""
    from paos_b200 import _lib, chain

    assert _lib.lib.paos_abi_struct_size(0) == C.sizeof(chain.Surface)
    assert _lib.lib.paos_abi_struct_size(1) == C.sizeof(chain.Snapshot)
    assert _lib.lib.paos_abi_struct_size(2) == C.sizeof(_lib.PaosStats)
    assert _lib.lib.paos_abi_struct_size(3) == C.sizeof(chain.ChainArgs)
    assert _lib.lib.paos_abi_struct_size(99) == -1


def test_build_record_matches_the_tree():
    """The loaded library says which source tree it was compiled from (paos_build_info); it must be this one."""
    import importlib.util

    from paos_b200 import _lib

    spec = importlib.util.spec_from_file_location("_paos_b200_build", os.path.join(ROOT, "paos_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    info = _lib.lib.paos_build_info().decode()
    assert mod.source_hash() in info, f"stale library: {info}"


def test_batch_entry_points_validate_their_arguments():
    from paos_b200 import _lib

    assert 2 <= _lib.lib.paos_batch_capacity() <= 64
    assert _lib.lib.paos_wfo_begin_record(None) == _lib.PAOS_ERR_ARG
    assert _lib.lib.paos_batch_execute(None, 1) == _lib.PAOS_ERR_ARG
    hs = (C.c_void_p * 2)()
    assert _lib.lib.paos_batch_execute(hs, 0) == _lib.PAOS_ERR_ARG
    assert _lib.lib.paos_batch_execute(hs, 2) == _lib.PAOS_ERR_ARG  # null handles
    assert _lib.lib.paos_batch_chain_run(hs, 2, None) == _lib.PAOS_ERR_ARG
    assert _lib.lib.paos_batch_execute(hs, _lib.lib.paos_batch_capacity() + 1) == _lib.PAOS_ERR_ARG
