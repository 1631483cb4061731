"""CPU-only checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/b200df.h declares, its structs have the documented layout, and -- there being no CPU fallback --
context creation fails loudly without a GPU.  No compute is attempted here."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200df.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bdf_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(rdf):
    lib = rdf.native.lib()
    declared = declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/b200df.h but not exported by libb200df.so"
    assert sorted(rdf.native.EXPORTED_SYMBOLS) == declared  # the Python binding covers the whole header
    out = subprocess.run(["nm", "-D", "--defined-only", rdf.native.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (bdf_[a-z0-9_]+)", out))
    assert exported == set(declared), exported ^ set(declared)


def test_header_compiles_as_c_and_struct_layout(tmp_path, rdf):
    """The header must be plain C (the Rust/cgo/JNI side reads it), and ctypes must agree on the layout."""
    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stdio.h>\n#include "b200df.h"\n'
        "int main(void){printf(\"%zu %zu %zu %zu\\n\", sizeof(bdf_view), sizeof(bdf_out), sizeof(bdf_agg4), sizeof(bdf_launch_record));return 0;}\n")
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    N = rdf.native
    assert sizes == [C.sizeof(N.View), C.sizeof(N.Out), C.sizeof(N.Agg4), C.sizeof(N.LaunchRecord)]


def test_cubin_is_sm_100a(rdf):
    out = subprocess.run(["cuobjdump", "-lelf", rdf.native.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out[:400]


def test_no_gpu_means_loud_failure(rdf):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(rdf.ArrowError) as ei:
        rdf.Context(0)
    assert "no CPU fallback" in str(ei.value)
    # and the operator API does not quietly compute on the host either
    a = rdf.PrimitiveArray.from_pylist(rdf.I32, [1, 2, 3])
    with pytest.raises(rdf.ArrowError):
        rdf.ScalarFunctions.add([a], [a])


def test_product_does_not_touch_the_oracle():
    """oracle/ is test infrastructure: nothing under rust-dataframe_b200/ may import, link or call it."""
    pkg = os.path.join(ROOT, "rust-dataframe_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.path.basename(dirpath) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".cpp", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert "liboracle" not in text and "pyoracle" not in text and "oracle.h" not in text, os.path.join(dirpath, f)
    deps = subprocess.run(["ldd", os.path.join(pkg, "libb200df.so")], capture_output=True, text=True).stdout
    assert "oracle" not in deps
