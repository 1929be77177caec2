"""CPU checks of the drop-in boundary: the shared library loads, exports every symbol include/ruviii.h declares,
the ctypes structs match the C structs, and every compute entry point fails loudly without a GPU."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from ruviii_b200 import _cabi

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
HEADER = os.path.join(ROOT, "include", "ruviii.h")


def _header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ruviii_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    syms = _header_symbols()
    assert len(syms) >= 16
    assert sorted(_cabi.SYMBOLS) == syms
    lib = _cabi.load_library()
    for s in syms:
        assert hasattr(lib, s), s
    exported = subprocess.run(["nm", "-D", "--defined-only", _cabi.LIB_PATH], capture_output=True, text=True).stdout
    names = {line.split()[-1] for line in exported.splitlines() if " T " in line}
    assert set(syms) <= names
    assert all(n.startswith("ruviii_") for n in names), "only the C ABI may be exported: %s" % sorted(names)[:5]


def test_struct_layouts_match_header(tmp_path):
    """Compile a tiny C program against the header and compare sizeof/offsetof with the ctypes mirrors."""
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ruviii.h"\nint main(void){'
                   'printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(ruviii_params), offsetof(ruviii_params, pseudo_count),'
                   'offsetof(ruviii_params, want_ps_rows), sizeof(ruviii_info), offsetof(ruviii_info, ms_log),'
                   'offsetof(ruviii_info, eig_top)); return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(_cabi.Params), _cabi.Params.pseudo_count.offset, _cabi.Params.want_ps_rows.offset,
            C.sizeof(_cabi.Info), _cabi.Info.ms_log.offset, _cabi.Info.eig_top.offset]
    assert got == want


def test_abi_version_and_no_gpu_behaviour():
    lib = _cabi.load_library()
    assert lib.ruviii_abi_version() == 1
    if lib.ruviii_device_count() > 0:
        pytest.skip("a GPU is present; the loud-failure path is exercised on CPU boxes")
    with pytest.raises(_cabi.RuviiiError) as e:
        _cabi.Engine(devices=[0])
    assert e.value.code == _cabi.ERR_CUDA and "no CPU fallback" in str(e.value)
    with pytest.raises(_cabi.RuviiiError):
        _cabi.Engine(rank=0, world=1, device=0)


def test_missing_library_is_loud(monkeypatch):
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", "/nonexistent/libruviii_b200.so")
    with pytest.raises(RuntimeError, match="no CPU or PyTorch fallback"):
        _cabi.load_library()


def test_product_never_imports_oracle():
    """The shipped path may not import, link or execute anything under oracle/ (task rule 3)."""
    pkg = os.path.join(ROOT, "ruviii_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert not re.search(r"^\s*(from|import)\s+oracle", text, flags=re.M), os.path.join(dirpath, f)
                assert "ruviii_oracle" not in text, os.path.join(dirpath, f)
