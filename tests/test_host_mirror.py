"""The C++ host mirror (rust-dataframe_b200/host/) of the reference's operator interface: builds against the
C ABI on CPU; on the GPU it runs the reference's own unit tests re-stated in C++ (test_reference.cpp)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "rust-dataframe_b200", "host")


def test_host_mirror_builds_and_links(rdf):
    subprocess.check_call(["make", "-C", HOST], stdout=subprocess.DEVNULL)
    exe = os.path.join(HOST, "test_reference")
    assert os.path.exists(exe)
    deps = subprocess.run(["ldd", exe], capture_output=True, text=True).stdout
    assert "libb200df.so" in deps and "not found" not in deps and "oracle" not in deps


def test_ipc_entries_from_cpp(rdf, tmp_path):
    """DataFrame::from_arrow / to_arrow through the C ABI from C++ (host/ipc.hpp): needs no GPU."""
    subprocess.check_call(["make", "-C", HOST, "test_ipc_host"], stdout=subprocess.DEVNULL)
    r = subprocess.run([os.path.join(HOST, "test_ipc_host"), str(tmp_path / "cpp.arrow")], capture_output=True, text=True, timeout=120)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0 and "IPC HOST OK" in r.stdout


@pytest.mark.gpu
def test_reference_unit_tests_in_cpp(rdf):
    subprocess.check_call(["make", "-C", HOST], stdout=subprocess.DEVNULL)
    r = subprocess.run([os.path.join(HOST, "test_reference")], capture_output=True, text=True, timeout=300)
    print(r.stdout[-3000:], r.stderr[-2000:])
    assert r.returncode == 0 and "ALL OK" in r.stdout
