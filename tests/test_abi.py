"""CPU-only checks of the drop-in boundary: the C-ABI shared library builds for sm_100a, loads,
and exports every symbol include/greyjack_b200.h declares.  No compute call is made here (there
is no GPU in the build container and the engine has no CPU fallback -- which is also checked)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import greyjack_b200 as gj
from greyjack_b200 import _lib, instances as inst

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "greyjack_b200.h")


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(gj.LIB_PATH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "greyjack-solver-rust_b200"), "-s",
                               "libgreyjack_b200.so"])
    return _lib.load()


def header_symbols():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"GJ_API\s+[\w\s\*]+?\b(gj_\w+)\s*\(", text)))


def test_header_declares_what_the_binding_lists():
    assert header_symbols() == sorted(_lib.EXPORTED)


def test_library_exports_every_declared_symbol(lib):
    missing = [s for s in header_symbols() if not hasattr(lib, s)]
    assert not missing, f"not exported: {missing}"


def test_abi_version_and_struct_sizes(lib):
    assert lib.gj_abi_version() >= 1
    # the ctypes mirrors must match the C layout (checked against the sizes the library reports)
    lib.gj_sizeof_problem_desc.restype = C.c_size_t
    lib.gj_sizeof_agent_params.restype = C.c_size_t
    assert lib.gj_sizeof_problem_desc() == C.sizeof(_lib.ProblemDesc)
    assert lib.gj_sizeof_agent_params() == C.sizeof(_lib.AgentParams)


def test_no_cpu_fallback(lib):
    """Without a CUDA device every compute entry point must fail loudly (GJ_ERR_CUDA)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert lib.gj_device_count() == 0
    spec = inst.tsp(16, seed=1)
    with pytest.raises(gj.GjError) as e:
        gj.Problem(spec)
    assert "no CUDA device" in str(e.value) or "status 2" in str(e.value)


def test_bad_arguments_do_not_crash(lib):
    assert lib.gj_problem_create(None, 0, None) != 0
    assert b"null" in lib.gj_last_error()
    assert lib.gj_problem_levels(None) == 0
    assert lib.gj_score_plain(None, None, C.c_int64(0), None) != 0
    assert lib.gj_islands_step(None, C.c_int64(1), None) != 0
    lib.gj_problem_destroy(None)
    lib.gj_islands_destroy(None)


def test_sass_is_sm100a_only():
    """The shipped library carries sm_100a code and nothing else (no multi-arch dispatch)."""
    out = subprocess.run(["cuobjdump", "-lelf", gj.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+\w?)", out))
    assert archs == {"100a"}, archs
