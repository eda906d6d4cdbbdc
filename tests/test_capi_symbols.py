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


def test_index_struct_mirror_matches_the_header(tmp_path):
    """ctypes mirror of `stb200_index` == the C struct in include/stb200.h (size and every field offset), checked by
    compiling a tiny C program against the header with the host compiler."""
    import shutil
    import subprocess
    import pytest
    from stratified_transformer_b200 import _cabi
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        pytest.skip("no host C compiler")
    fields = [name for name, _ in _cabi.IndexStruct._fields_]
    src = tmp_path / "layout.c"
    src.write_text('#include <stddef.h>\n#include <stdio.h>\n#include "stb200.h"\nint main(void) {\n'
                   '  printf("%zu\\n", sizeof(stb200_index));\n' +
                   "".join(f'  printf("%zu\\n", offsetof(stb200_index, {f}));\n' for f in fields) + "  return 0;\n}\n")
    exe = tmp_path / "layout"
    subprocess.run([cc, "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert out[0] == ctypes.sizeof(_cabi.IndexStruct)
    assert out[1:] == [getattr(_cabi.IndexStruct, f).offset for f in fields]
