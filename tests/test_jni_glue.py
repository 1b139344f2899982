"""The JNI glue a maintainer adds on the reference side compiles warning-free against the JDK's
function signatures (a stand-in jni.h: no JDK in this image) and links against libifb200.so, so every ABI entry
point it names exists with the argument types it passes."""
import os
import re
import subprocess

from conftest import ROOT

JNI_C = os.path.join(ROOT, "isolation-forest_b200", "jvm", "ifb200_jni.c")
SCALA = os.path.join(ROOT, "isolation-forest_b200", "jvm", "NativeForest.scala")


def test_jni_glue_compiles_and_links(nat, tmp_path):
    so = tmp_path / "libifb200_jni.so"
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-shared", "-fPIC", "-I", os.path.join(ROOT, "tests", "jni_stub"),
           "-I", os.path.join(ROOT, "include"), JNI_C, "-L", os.path.dirname(nat.LIB_PATH), "-lifb200",
           "-Wl,--no-undefined", "-Wl,-rpath," + os.path.dirname(nat.LIB_PATH), "-o", str(so)]
    out = subprocess.run(cmd, capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    syms = subprocess.check_output(["nm", "-D", "--defined-only", str(so)]).decode()
    exported = set(re.findall(r"Java_com_linkedin_relevance_isolationforest_gpu_NativeForest_00024_(\w+)", syms))
    # every @native method of the Scala object has its C implementation, and vice versa
    natives = set(re.findall(r"@native\s+def\s+(\w+)", open(SCALA).read()))
    assert natives == exported, (sorted(natives - exported), sorted(exported - natives))
    assert {"exportTables", "createExtended", "scoreHost", "fitHost", "commInit"} <= exported


def test_jni_glue_never_holds_critical_regions():
    src = open(JNI_C).read()
    code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    assert "GetPrimitiveArrayCritical" not in code
