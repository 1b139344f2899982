"""CPU tests of the drop-in boundary: libifb200.so loads and exports exactly what include/ifb200.h declares."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT


def header_functions():
    src = open(os.path.join(ROOT, "include", "ifb200.h")).read()
    return sorted(set(re.findall(r"IFB_API\s+[\w\s\*]+?\b(ifb_\w+)\s*\(", src)))


def test_header_symbols_are_bound_and_exported(nat):
    names = header_functions()
    assert len(names) >= 20
    assert sorted(nat.SYMBOLS) == names, "python binding and header disagree"
    lib = nat.lib()
    for n in names:
        assert getattr(lib, n) is not None
    out = subprocess.check_output(["nm", "-D", "--defined-only", nat.LIB_PATH]).decode()
    exported = sorted(set(re.findall(r"\bT (ifb_\w+)", out)))
    assert exported == names, "exported symbol set differs from the header"


def test_library_is_sm100a_only(nat):
    out = subprocess.check_output(["cuobjdump", "-lelf", nat.LIB_PATH]).decode()
    archs = set(re.findall(r"sm_(\w+)\.cubin", out))
    assert archs == {"100a"}, archs


def test_no_compute_without_gpu_fails_loudly(nat):
    """On a CPU-only box the product path must refuse to run (no fallback)."""
    if nat.device_count() > 0:
        pytest.skip("GPU present")
    t = dict(extended=False, num_trees=1, num_samples=256, node_off=np.array([0, 1], np.int32),
             left=np.array([-1], np.int32), right=np.array([-1], np.int32), feature=np.array([-1], np.int32),
             threshold=np.zeros(1), num_instances=np.array([256], np.int64))
    with pytest.raises(RuntimeError, match="no CPU fallback|no CUDA device"):
        nat.NativeForest.from_tables(t)


def test_scalar_helpers(nat, oracle):
    assert nat.lib().ifb_abi_version() == 1
    for n in (0, 1, 2, 3, 10, 255, 256, 100000, 2**40, 2**63 - 1):
        assert np.float32(nat.lib().ifb_avg_path_length(n)) == oracle.avg_path_length(n)


def test_product_never_references_oracle():
    """The product tree must not import, link or open anything under oracle/."""
    pkg = os.path.join(ROOT, "isolation-forest_b200")
    for dp, _, fns in os.walk(pkg):
        for fn in fns:
            if fn.endswith((".py", ".cu", ".h", ".cpp", ".cc", "Makefile")):
                assert "oracle" not in open(os.path.join(dp, fn), errors="ignore").read().lower(), os.path.join(dp, fn)


def _build_c_consumer(tmp_path):
    exe = tmp_path / "abi_smoke"
    lib = os.path.join(ROOT, "isolation-forest_b200")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-o", str(exe), "-L" + lib, "-lifb200",
                           "-Wl,-rpath," + lib])
    return exe


def test_header_is_plain_c_and_links(nat, tmp_path):
    """include/ifb200.h compiles as strict C99 and the library links from C (what a JNI stub needs)."""
    exe = _build_c_consumer(tmp_path)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "abi_smoke ok" in out.stdout


@pytest.mark.gpu
def test_c_consumer_scores_on_gpu(nat, tmp_path):
    exe = _build_c_consumer(tmp_path)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "GPU: path lengths 4.7488804 6.1433091" in out.stdout, out.stdout + out.stderr
