"""The C-ABI library builds, loads and exports every symbol include/xdb200.h declares
(no compute calls: this runs without a GPU)."""
import ctypes
import os
import subprocess

import pytest

from xdiffusion_b200 import _lib


@pytest.fixture(scope="module")
def built():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return _lib.LIB_PATH


def test_header_parses():
    decls = _lib.parse_header()
    assert len(decls) >= 25
    for must in ("xd_gemm_bf16_tc", "xd_conv3x3_bf16_tc", "xd_attention_bf16", "xd_groupnorm_apply",
                 "xd_layernorm_modulate", "xd_timestep_embed", "xd_patchify", "xd_unpatchify", "xd_cfg_combine",
                 "xd_sampler_step", "xd_schedule_advance"):
        assert must in decls, must
    assert decls["xd_gemm_bf16_tc"][1].count(ctypes.c_void_p) == 8


def test_library_exports_every_declared_symbol(built):
    l = ctypes.CDLL(built)
    for name in _lib.parse_header():
        assert hasattr(l, name), name
    assert _lib.lib().xd_abi_version() == 3


def test_sass_is_blackwell_native(built):
    """tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM, TMA -> UTMALDG (B200_PROFILING.md)."""
    sass = subprocess.run(["cuobjdump", "-sass", built], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    for mnemonic in ("UTCHMMA", "LDTM", "UTMALDG"):
        assert mnemonic in sass, mnemonic


def test_ops_register_and_reject_cpu_tensors(built):
    import torch

    from xdiffusion_b200 import ops  # noqa: F401
    a = torch.randn(128, 64).bfloat16()
    with pytest.raises((NotImplementedError, RuntimeError)):
        torch.ops.xdb200.gemm(a, None, a, None, 0, None, 1, None, torch.empty(128, 128), 0)
