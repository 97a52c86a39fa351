"""The drop-in boundary is the C ABI of vgsim_b200/libvgsim_b200.so: every function include/vgsim_b200.h
declares must be exported by the built library with C linkage, the ctypes binding must cover all of
them, and no torch/CUDA type may leak into a signature.  No compute calls here (no GPU needed)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "vgsim_b200.h")


def declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vgsim_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_path():
    names = declared()
    for must in ("vgsim_create", "vgsim_upload_params", "vgsim_simulate_direct", "vgsim_simulate_tau", "vgsim_genealogy",
                 "vgsim_propensities", "vgsim_get_event_log", "vgsim_get_tree", "vgsim_get_mutations", "vgsim_get_migrations"):
        assert must in names
    assert len(names) >= 35


def test_library_exports_every_declared_symbol():
    from vgsim_b200 import _capi
    lib = ctypes.CDLL(_capi.LIB_PATH)
    missing = [n for n in declared() if not hasattr(lib, n)]
    assert not missing, missing
    # the Python binding declares argtypes for exactly the header's functions
    assert sorted(_capi.EXPORTS) == declared()


def test_header_is_plain_c():
    src = open(HEADER).read()
    assert 'extern "C"' in src
    body = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for banned in ("torch", "at::", "cudaStream_t", "std::", "Tensor"):
        assert banned not in body
    # compiles as C (gcc, no CUDA headers needed)
    r = subprocess.run(["gcc", "-std=c99", "-fsyntax-only", "-x", "c", HEADER], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout


def test_version_and_error_string_without_gpu():
    from vgsim_b200 import _capi
    assert _capi.lib.vgsim_version() >= 100
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(_capi.VgsimError, match="no CUDA device"):
            _capi.Handle(0, 1, 1)
