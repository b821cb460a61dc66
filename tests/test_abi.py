"""The C-ABI boundary: include/deepatlas_b200.h, the ctypes table in deepatlas_b200/_lib.py and the symbols the
built shared library exports must agree (no compute calls -- runs without a GPU)."""
import ctypes
import os
import subprocess

import pytest

from abi_util import parse_header


def test_header_matches_ctypes_table():
    from deepatlas_b200 import _lib
    hdr = parse_header()
    assert set(hdr) == set(_lib.SIGNATURES), (set(hdr) ^ set(_lib.SIGNATURES))
    ret_map = {"int": {"rc", "int"}, "int64_t": {"size"}, "const char*": {"str"}}
    for name, (codes, ret) in hdr.items():
        py_codes, py_ret = _lib.SIGNATURES[name]
        assert codes == py_codes, f"{name}: header {codes} vs ctypes {py_codes}"
        assert py_ret in ret_map[ret], f"{name}: return kind {py_ret} vs header {ret}"


def test_library_exports_every_declared_symbol(built_lib):
    from deepatlas_b200 import _lib
    hdr = parse_header()
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {line.split()[-1] for line in out.splitlines() if " T " in line}
    missing = set(hdr) - exported
    assert not missing, f"declared in the header but not exported: {sorted(missing)}"
    stray = {s for s in exported if s.startswith("da_")} - set(hdr)
    assert not stray, f"exported but not declared in include/deepatlas_b200.h: {sorted(stray)}"
    assert built_lib.da_version() >= 100
    assert isinstance(built_lib.da_last_error(), bytes)


def test_size_queries_without_gpu(built_lib):
    """Pure host arithmetic entry points work without a device."""
    from deepatlas_b200 import _lib
    assert _lib.size("da_conv3d_pack_bytes", 48, 16, 3) >= 48 * 27 * 16 * 4
    assert _lib.size("da_conv3d_wgrad_workspace_bytes", 48, 16, 3) >= 48 * 16 * 27 * 4
    assert _lib.size("da_dice_workspace_bytes", 1, 32, 4915200) > 0
    assert _lib.size("da_bn_workspace_bytes", 64) > 0
    assert _lib.size("da_channel_sum_workspace_bytes", 16) > 0
    assert _lib.size("da_lncc_coef_bytes", 1, 160, 192, 160, 9, 1) > 0


def test_no_cpu_fallback_and_missing_library_is_loud(monkeypatch, tmp_path):
    import torch
    from deepatlas_b200 import _lib, ops
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.conv3d(torch.zeros(1, 1, 4, 4, 4), torch.zeros(2, 1, 3, 3, 3))
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.warp3d(torch.zeros(1, 1, 4, 4, 4), torch.zeros(1, 3, 4, 4, 4))
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "libdeepatlas_b200.so"))
    with pytest.raises(RuntimeError, match="not built"):
        _lib.load()


def test_header_is_plain_c(tmp_path):
    """include/deepatlas_b200.h is what a C (not C++) host would include: it must compile as C99 with warnings on."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    src = tmp_path / "abi.c"
    src.write_text('#include "deepatlas_b200.h"\nint main(void) { return da_version() == 0; }\n')
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-I", inc, str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
