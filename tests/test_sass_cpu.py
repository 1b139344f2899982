"""The shipped libifb200.so really carries the Blackwell instructions DESIGN.md describes (no GPU needed: cuobjdump
disassembles the sm_100a cubin).  Guards against a build that silently lost a kernel family or fell back to a generic
code path."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "isolation-forest_b200", "libifb200.so")


@pytest.fixture(scope="module")
def sass():
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    if not os.path.exists(SO):
        pytest.fail("libifb200.so is not built (run __graft_entry__.build())")
    return subprocess.check_output(["cuobjdump", "-sass", SO]).decode()


def test_library_is_sm_100a_only(sass):
    archs = set(re.findall(r"arch = (sm_\d+a?)", sass))
    assert archs == {"sm_100a"}, archs


@pytest.mark.parametrize("mnemonic,what", [
    ("UTCHMMA ", "tcgen05.mma (cta_group::1) of the extended-forest GEMM"),
    ("UTCHMMA.2CTA", "tcgen05.mma.cta_group::2 of the opt-in CTA-pair variant"),
    ("LDTM.x32", "tcgen05.ld of the accumulator drain"),
    ("UTMALDG.2D ", "TMA tile loads (standard kernel row tiles, GEMM operands)"),
    ("UTMALDG.2D.MULTICAST", "TMA multicast of the hyperplane tiles inside a cluster"),
    ("UTMALDG.2D.2CTA", "TMA loads of the CTA-pair variant"),
    ("UTCBAR", "tcgen05.commit"),
    ("UTMAPF.L2.2D", "L2 prefetch of the next row tile (256-row standard tiles)"),
    ("UBLKCP.S.G", "cp.async.bulk (block descriptors, rank-kernel feature columns)"),
    ("SYNCS.PHASECHK.TRANS64.TRYWAIT", "mbarrier waits"),
    ("FMNMX3", "3-input abs-min of the drain's chunk-wide ambiguity test"),
])
def test_instruction_is_present(sass, mnemonic, what):
    assert mnemonic in sass, f"{mnemonic.strip()} missing: {what}"


@pytest.mark.parametrize("kernel", ["score_std_kernel", "score_std_rank_kernel", "score_ext_tc_kernel", "score_ext_dense_kernel",
                                    "score_ext_wide_kernel", "fit_kernel", "ext_tc_prepare_rows", "ext_tc_prepare_cols"])
def test_kernel_family_is_compiled(sass, kernel):
    assert re.search(r"Function : \S*" + kernel, sass), kernel
