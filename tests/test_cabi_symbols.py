"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/*.h declares."""
import ctypes
import os
import re

import pytest

from afldm_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "afldm_b200.h")).read()
    return sorted(set(re.findall(r"AFLDM_API[^;(]*?\b(afldm_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.LIB_PATH


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    assert len(syms) >= 20
    for must in ("afldm_filtered_act_f32", "afldm_up2_ideal_f32", "afldm_lpf_down2_f32", "afldm_conv2d_f32",
                 "afldm_attention_f32", "afldm_groupnorm_affine_f32", "afldm_upfirdn2d_f32"):
        assert must in syms


def test_library_exports_every_declared_symbol(built):
    handle = ctypes.CDLL(built)
    for name in declared_symbols():
        assert hasattr(handle, name), f"{name} declared in include/afldm_b200.h but not exported"


def test_python_binding_covers_header(built):
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.lib()
    assert lib.afldm_abi_version() == 1
    assert lib.afldm_error_string(-3).decode().startswith("afldm")
    assert isinstance(_lib.launch_count(), int)


def test_argument_checks_do_not_need_a_gpu(built):
    """Bad arguments are rejected on the host before any launch (returns AFLDM_E_*)."""
    lib = _lib.lib()
    assert lib.afldm_filtered_act_f32(None, None, 1, 8, 8, 32, 1, None, None, None, 0, None) == -2
    assert lib.afldm_resample_workspace_floats(0, 2, 32, 32, 64) == 0
    assert lib.afldm_resample_workspace_floats(0, 2, 64, 64, 64) == 2 * 2 * 64 * 128 * 64
    assert lib.afldm_conv2d_f32(None, 4, None, None, None, 0, None, 0, None, 4, 1, 2, 2, 4, 4, 3, 0, None, 0, None, None) == -2
    assert lib.afldm_groupnorm_scratch_floats(16, 1024, 192) == 0
    assert lib.afldm_conv2d_workspace_floats(16, 32, 32, 192, 192, 3, 0) == 0      # 128 x 3 tiles: no split-K
    assert lib.afldm_conv2d_workspace_floats(16, 2, 2, 1536, 768, 3, 0) > 0        # 2x2 level: split-K


def test_sass_is_sm100a_only(built):
    out = os.popen(f"cuobjdump -lelf {built} 2>/dev/null").read()
    if not out.strip():
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
