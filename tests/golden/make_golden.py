#!/usr/bin/env python
"""Regenerate tests/golden/*.npz from the reference checkout (run in the build container only).

/root/reference does not exist on the GPU box, so everything the parity tests need from the reference's
own fixtures is distilled here into small numpy archives:

  mammography.npz   X (11183 x 6 f32 = `.toFloat` of the CSV doubles), label      <- IFR/mammography.csv
  shuttle.npz       X (49097 x 9 f32), label                                        <- IFR/shuttle.csv
  model_<name>.npz  pre-order node tables + metadata JSON of the four saved models  <- IFR/saved*Model,
                    ONNX/test/resources/savedIsolationForestModel/{mammography,shuttle}Model
  mammography_scores.npz  the 11,183 reference-computed f64 scores + predicted labels
                    <- ONNX/test/resources/savedIsolationForestModel/mammographyModel/mammographyOutlierScores.csv
  tree_text.json    sha256 + length of IFR/expectedTreeStructure.txt / expectedExtendedTreeStructure.txt

(IFR = isolation-forest/src/test/resources, ONNX = isolation-forest-onnx.)  The Avro files are decoded by
oracle/avro_min.py; nothing here is produced by product code.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import avro_min  # noqa: E402

REF = os.environ.get("IFB_REFERENCE", "/root/reference")
IFR = os.path.join(REF, "isolation-forest/src/test/resources")
ONNX = os.path.join(REF, "isolation-forest-onnx/test/resources")

MODELS = {
    "std_mammography_spark23": os.path.join(IFR, "savedIsolationForestModel"),
    "ext_mammography": os.path.join(IFR, "savedExtendedIsolationForestModel"),
    "std_mammography_onnx": os.path.join(ONNX, "savedIsolationForestModel/mammographyModel"),
    "std_shuttle_onnx": os.path.join(ONNX, "savedIsolationForestModel/shuttleModel"),
}


def main():
    for name in ("mammography", "shuttle"):
        a = np.loadtxt(os.path.join(IFR, name + ".csv"), delimiter=",", comments="#")
        np.savez_compressed(os.path.join(HERE, name + ".npz"), X=a[:, :-1].astype(np.float32),
                            label=a[:, -1].astype(np.uint8))
        print(name, a.shape)
    for name, path in MODELS.items():
        meta, recs, codec = avro_min.read_model_dir(path)
        t = avro_min.forest_arrays(meta, recs)
        arrays = {k: v for k, v in t.items() if isinstance(v, np.ndarray)}
        scalars = {k: v for k, v in t.items() if not isinstance(v, np.ndarray)}
        np.savez_compressed(os.path.join(HERE, f"model_{name}.npz"), metadata_json=np.array(json.dumps(meta)),
                            scalars_json=np.array(json.dumps(scalars)), codec=np.array(codec), **arrays)
        print(name, codec, scalars)
    g = np.loadtxt(os.path.join(ONNX, "savedIsolationForestModel/mammographyModel/mammographyOutlierScores.csv"),
                   delimiter=",", skiprows=1)
    assert (g[:, 0] == np.arange(len(g))).all()
    np.savez_compressed(os.path.join(HERE, "mammography_scores.npz"), score=g[:, 1].astype(np.float64),
                        predicted=g[:, 2].astype(np.uint8), label=g[:, 3].astype(np.uint8),
                        X=g[:, 4:].astype(np.float32))
    texts = {}
    for fn in ("expectedTreeStructure.txt", "expectedExtendedTreeStructure.txt"):
        b = open(os.path.join(IFR, fn), "rb").read()
        texts[fn] = {"sha256": hashlib.sha256(b.strip()).hexdigest(), "length": len(b.strip())}
    json.dump(texts, open(os.path.join(HERE, "tree_text.json"), "w"), indent=1)
    print(texts)


if __name__ == "__main__":
    main()
