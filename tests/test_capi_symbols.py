"""The C-ABI library loads on a machine without a GPU and exports every symbol include/stb200.h declares
(no compute calls here).  Also: the product package never imports the oracle."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "stb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(stb200_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from stratified_transformer_b200 import _cabi
    lib = _cabi.load()
    names = declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/stb200.h but not exported by libstb200.so"
    # and the ctypes table binds exactly the declared functions
    assert set(_cabi.EXPORTED_SYMBOLS) == set(names), set(_cabi.EXPORTED_SYMBOLS) ^ set(names)


def test_error_reporting_without_gpu_or_with_bad_args():
    from stratified_transformer_b200 import _cabi
    lib = _cabi.load()
    rc = lib.stb200_attention_step1_forward_v2(4, 4, 2, 48, 0, None, None, None, None, None, None)   # C/h = 24
    assert rc == 1 and b"d != 16" in lib.stb200_last_error()
    rc = lib.stb200_attention_step1_forward_v2(4, 4, 3, 48, 0, None, None, None, None, None, None)   # null pointers
    assert rc == 2
    assert lib.stb200_version() >= 100 and lib.stb200_launch_count() >= 0


def test_product_package_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "stratified_transformer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"(^|\s)(from|import)\s+oracle|oracle[/.](_ref|_build|attention_oracle|index_oracle|fps_oracle|ref_cuda)", text), \
                    f"{f} references the oracle"


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from stratified_transformer_b200 import _cabi
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", str(tmp_path / "nope.so"))
    import pytest
    with pytest.raises(_cabi.Stb200Error, match="no CPU fallback"):
        _cabi.load()
