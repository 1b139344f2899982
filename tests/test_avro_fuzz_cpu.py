"""Corrupted / truncated model files must surface as exceptions from the native Avro + metadata reader, never as a
crash or a hang (run in a child process so that a crash would be seen as a failed test, not a dead pytest)."""
import os
import subprocess
import sys

import pytest

from conftest import REFERENCE, ROOT

CHILD = r'''
import glob, os, random, shutil, sys
sys.path.insert(0, sys.argv[1])
import numpy as np
import __graft_entry__ as g
E = g.load_package().estimators
src, work, seed = sys.argv[2], sys.argv[3], int(sys.argv[4])
if src == "self":                       # a model written by this repo's own writer (deflate codec)
    z = np.load(os.path.join(sys.argv[1], "tests", "golden", "model_std_mammography_onnx.npz"), allow_pickle=True)
    t = {k: z[k] for k in ("node_off", "left", "right", "feature", "threshold", "num_instances")}
    t["num_trees"] = len(t["node_off"]) - 1
    src = work + "_src"
    E.IsolationForestModel.from_tables("fuzz", t, 256, 6, 6).write().overwrite().save(src)
random.seed(seed)
outcomes = {"loaded": 0, "error": 0}
for it in range(24):
    shutil.rmtree(work, ignore_errors=True)
    shutil.copytree(src, work)
    f = random.choice(glob.glob(work + "/data/*.avro") + glob.glob(work + "/metadata/part-*"))
    b = bytearray(open(f, "rb").read())
    mode = it % 3
    if mode == 0:
        b = b[: random.randrange(0, len(b))]
    elif mode == 1:
        for _ in range(random.randrange(1, 6)):
            b[random.randrange(len(b))] = random.randrange(256)
    else:
        i = random.randrange(len(b))
        b[i:i] = bytes(random.randrange(256) for _ in range(random.randrange(1, 9)))
    open(f, "wb").write(bytes(b))
    try:
        E.IsolationForestModel.load(work)
        outcomes["loaded"] += 1           # e.g. a flipped bit inside a float: still a well-formed file
    except (ValueError, RuntimeError):
        outcomes["error"] += 1
print("fuzz ok", outcomes)
'''


def _run(src, tmp_path, seed):
    out = subprocess.run([sys.executable, "-c", CHILD, ROOT, src, str(tmp_path / "work"), str(seed)], capture_output=True,
                         text=True, timeout=300)
    assert out.returncode == 0, (out.returncode, out.stdout[-500:], out.stderr[-1500:])
    assert "fuzz ok" in out.stdout and "'error': 0" not in out.stdout


def test_own_writer_files_fuzzed(tmp_path):
    _run("self", tmp_path, 11)


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference checkout not mounted")
def test_reference_written_files_fuzzed(tmp_path):
    _run(os.path.join(REFERENCE, "isolation-forest/src/test/resources/savedIsolationForestModel"), tmp_path, 12)
