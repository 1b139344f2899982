"""a model trained and written by this engine must be consumable by the reference's ONNX converter
(isolation-forest-onnx/src/isolationforestonnx/isolation_forest_converter.py), which reads the same metadata JSON +
Avro node rows.  The converter needs `avro`, `onnx` and `onnxruntime`, none of which exist in this image (no network),
so the test skips here; it documents and pins the contract for an environment that has them."""
import os
import sys

import numpy as np
import pytest

from conftest import REFERENCE

pytestmark = pytest.mark.gpu


def test_gpu_trained_model_converts_to_onnx(pkg, golden, tmp_path):
    pytest.importorskip("onnx")
    ort = pytest.importorskip("onnxruntime")
    pytest.importorskip("avro")
    src = os.path.join(REFERENCE, "isolation-forest-onnx", "src")
    if not os.path.isdir(src):
        pytest.skip("reference checkout not present")
    sys.path.insert(0, src)
    from isolationforestonnx.isolation_forest_converter import IsolationForestConverter

    X = golden.mammography["X"].astype(np.float64)
    model = pkg.IsolationForest().setNumEstimators(100).setContamination(0.02).setRandomSeed(1).fit(X)
    model.save(tmp_path / "m")
    data = [f for f in os.listdir(tmp_path / "m" / "data") if f.endswith(".avro")][0]
    conv = IsolationForestConverter(str(tmp_path / "m" / "data" / data), str(tmp_path / "m" / "metadata" / "part-00000"))
    onx = conv.convert()
    sess = ort.InferenceSession(onx.SerializeToString())
    got = sess.run(None, {"features": X.astype(np.float32)})[0].ravel()
    want = model.transform(X).outlierScore
    assert np.max(np.abs(got - want)) < 1e-5      # the reference's own Spark-vs-ONNX tolerance
