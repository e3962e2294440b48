"""The C-ABI libraries load without a GPU and export every symbol include/*.h declares; struct
layouts of the ctypes mirror agree with the C headers (checked by compiling a tiny C probe)."""
import ctypes as C
import os
import re
import subprocess

import pytest

import kiraray_b200 as krr
from kiraray_b200 import binding as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(krr_[a-z0-9_]+)\s*\(", text)))


def test_wfpt_library_exports_every_declared_symbol():
    lib = krr.load_wfpt()
    names = declared_functions("krr_wfpt.h")
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"libkrr_wfpt.so does not export {n}"
    assert lib.krr_wfpt_abi_version() == 6


def test_host_library_exports_every_declared_symbol():
    lib = krr.load_host()
    for n in declared_functions("krr_host_c.h"):
        assert hasattr(lib, n), f"libkrr_host.so does not export {n}"


def test_only_the_c_abi_is_exported():
    """-fvisibility=hidden: nothing but krr_* (and toolchain symbols) leaves libkrr_wfpt.so."""
    out = subprocess.run(["nm", "-D", "--defined-only", os.path.join(krr.lib_dir(), "libkrr_wfpt.so")], capture_output=True, text=True).stdout
    syms = [l.split()[-1] for l in out.splitlines() if " T " in l]
    leaked = [s for s in syms if not s.startswith("krr_") and not s.startswith("_")]
    assert not leaked, leaked[:10]


def test_ctypes_struct_layout_matches_header(tmp_path):
    names = ["KrrTextureDesc", "KrrSpectrumDesc", "KrrMaterialDesc", "KrrMeshDesc", "KrrSRT", "KrrTransformNodeDesc", "KrrInstanceDesc", "KrrLightDesc",
             "KrrMediumDesc", "KrrSceneOptions", "KrrSceneDesc", "KrrCameraData", "KrrColorSpaceData", "KrrStats"]
    src = tmp_path / "probe.c"
    src.write_text('#include <stdio.h>\n#include "krr_wfpt.h"\nint main(void){' +
                   "".join(f'printf("{n} %zu\\n", sizeof({n}));' for n in names) + "return 0;}\n")
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for n in names:
        assert C.sizeof(getattr(B, n)) == int(out[n]), n


def test_error_returns_instead_of_exit():
    """The reference's Log(Fatal)/CUDA_CHECK exit(1) (src/core/logger.cpp:100) become error codes."""
    lib = krr.load_wfpt()
    assert lib.krr_wfpt_create(b"{}", None) == -1  # KRR_E_INVALID
    assert b"null" in lib.krr_wfpt_last_error()
    assert lib.krr_wfpt_set_params(None, b"{}") == -1
    assert lib.krr_wfpt_render(None, None, None) == -1
    assert lib.krr_wfpt_resize(None, 0, 0) == -1


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    """No CPU fallback: a missing native library raises instead of silently degrading."""
    monkeypatch.setattr(B, "lib_dir", lambda: str(tmp_path))
    monkeypatch.setattr(B, "_wfpt", None)
    with pytest.raises(B.NativeLibraryMissing):
        B.load_wfpt()
