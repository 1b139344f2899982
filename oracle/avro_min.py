"""Minimal Avro object-container reader/writer (TEST INFRASTRUCTURE, not product code).

Used only by ``tests/`` and ``tests/golden/make_golden.py`` to read the reference's saved-model
fixtures (``data/*.avro`` written by spark-avro; see the reference
``IsolationForestModelReadWrite.scala:210-250`` and
``extended/ExtendedIsolationForestModelReadWrite.scala:216-263``) without a JVM, and to cross-check the
product's native C++ reader/writer.  Implements the public Avro 1.x container spec: magic ``Obj\\x01``,
metadata map, 16-byte sync marker, blocks of (count, byteSize, payload, sync); codecs ``null``,
``deflate`` (raw DEFLATE) and ``snappy`` (raw snappy block + 4-byte big-endian CRC32 of the
uncompressed bytes).  Decoding is schema-driven (record / union / array / int / long / float / double /
string / bytes / boolean / null).
"""
from __future__ import annotations

import json
import os
import struct
import zlib

MAGIC = b"Obj\x01"


class _Buf:
    __slots__ = ("b", "i")

    def __init__(self, b: bytes, i: int = 0):
        self.b = b
        self.i = i

    def read(self, n: int) -> bytes:
        v = self.b[self.i:self.i + n]
        if len(v) != n:
            raise EOFError("truncated avro data")
        self.i += n
        return v

    def long(self) -> int:
        shift = 0
        acc = 0
        while True:
            c = self.b[self.i]
            self.i += 1
            acc |= (c & 0x7F) << shift
            if not (c & 0x80):
                break
            shift += 7
        return (acc >> 1) ^ -(acc & 1)

    def eof(self) -> bool:
        return self.i >= len(self.b)


def snappy_decompress(src: bytes) -> bytes:
    """Raw snappy block format (public format description): varint length, then tagged elements."""
    i = 0
    n = 0
    shift = 0
    while True:
        c = src[i]
        i += 1
        n |= (c & 0x7F) << shift
        if not (c & 0x80):
            break
        shift += 7
    out = bytearray()
    L = len(src)
    while i < L:
        tag = src[i]
        i += 1
        t = tag & 3
        if t == 0:  # literal
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(src[i:i + nb], "little")
                i += nb
            ln += 1
            out += src[i:i + ln]
            i += ln
            continue
        if t == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | src[i]
            i += 1
        elif t == 2:
            ln = (tag >> 2) + 1
            off = src[i] | (src[i + 1] << 8)
            i += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(src[i:i + 4], "little")
            i += 4
        if off == 0 or off > len(out):
            raise ValueError("bad snappy copy offset")
        start = len(out) - off
        if off >= ln:
            out += out[start:start + ln]
        else:  # overlapping copy
            for k in range(ln):
                out.append(out[start + k])
    if len(out) != n:
        raise ValueError("snappy length mismatch")
    return bytes(out)


def _decode(schema, buf: _Buf):
    if isinstance(schema, list):  # union
        return _decode(schema[buf.long()], buf)
    if isinstance(schema, dict):
        t = schema["type"]
        if t == "record":
            return {f["name"]: _decode(f["type"], buf) for f in schema["fields"]}
        if t == "array":
            items = []
            while True:
                cnt = buf.long()
                if cnt == 0:
                    break
                if cnt < 0:
                    cnt = -cnt
                    buf.long()  # block byte size
                for _ in range(cnt):
                    items.append(_decode(schema["items"], buf))
            return items
        return _decode(t, buf)
    if schema in ("int", "long"):
        return buf.long()
    if schema == "double":
        return struct.unpack("<d", buf.read(8))[0]
    if schema == "float":
        return struct.unpack("<f", buf.read(4))[0]
    if schema == "null":
        return None
    if schema == "boolean":
        return buf.read(1) != b"\x00"
    if schema in ("string", "bytes"):
        v = buf.read(buf.long())
        return v.decode("utf-8") if schema == "string" else v
    raise ValueError(f"unsupported avro type {schema!r}")


def read_container(path: str):
    """Return (schema_dict, codec, [records]) of one Avro container file."""
    with open(path, "rb") as fh:
        buf = _Buf(fh.read())
    if buf.read(4) != MAGIC:
        raise ValueError("not an avro container")
    meta = {}
    while True:
        cnt = buf.long()
        if cnt == 0:
            break
        if cnt < 0:
            cnt = -cnt
            buf.long()
        for _ in range(cnt):
            k = buf.read(buf.long()).decode()
            meta[k] = buf.read(buf.long())
    sync = buf.read(16)
    schema = json.loads(meta["avro.schema"])
    codec = meta.get("avro.codec", b"null").decode()
    records = []
    while not buf.eof():
        cnt = buf.long()
        size = buf.long()
        payload = buf.read(size)
        if buf.read(16) != sync:
            raise ValueError("sync marker mismatch")
        if codec == "deflate":
            payload = zlib.decompress(payload, -15)
        elif codec == "snappy":
            crc = int.from_bytes(payload[-4:], "big")
            payload = snappy_decompress(payload[:-4])
            if zlib.crc32(payload) & 0xFFFFFFFF != crc:
                raise ValueError("snappy CRC mismatch")
        elif codec not in ("null", "uncompressed"):
            raise ValueError(f"unsupported codec {codec}")
        b2 = _Buf(payload)
        for _ in range(cnt):
            records.append(_decode(schema, b2))
    return schema, codec, records


def read_model_dir(path: str):
    """Read a saved model directory -> (metadata dict, [records], codec)."""
    with open(os.path.join(path, "metadata", "part-00000")) as fh:
        meta = json.loads(fh.readline())
    ddir = os.path.join(path, "data")
    recs = []
    codec = None
    for fn in sorted(os.listdir(ddir)):
        if fn.endswith(".avro"):
            _, codec, r = read_container(os.path.join(ddir, fn))
            recs += r
    return meta, recs, codec


def forest_arrays(meta: dict, recs: list):
    """Flatten decoded node rows into pre-order struct-of-arrays numpy tables (one per forest).

    Standard: node_off[T+1], left, right, feature (int32), threshold (f64), num_instances (int64).
    Extended: node_off, left, right, num_instances, offset (f64), hp_off[nodes+1] (int64),
              hp_idx (int32), hp_w (f32).
    """
    import numpy as np

    extended = "extendedNodeData" in recs[0]
    key = "extendedNodeData" if extended else "nodeData"
    trees = {}
    for r in recs:
        trees.setdefault(r["treeID"], []).append(r[key])
    T = len(trees)
    assert sorted(trees) == list(range(T))
    node_off = [0]
    left, right, ninst = [], [], []
    feat, thr = [], []
    off, hp_off, hp_idx, hp_w = [], [0], [], []
    for t in range(T):
        nodes = sorted(trees[t], key=lambda x: x["id"])
        assert [n["id"] for n in nodes] == list(range(len(nodes)))
        for n in nodes:
            left.append(n["leftChild"])
            right.append(n["rightChild"])
            ninst.append(n["numInstances"])
            if extended:
                off.append(n["offset"])
                hp_idx += n["indices"]
                hp_w += n["weights"]
                hp_off.append(len(hp_idx))
            else:
                feat.append(n["splitAttribute"])
                thr.append(n["splitValue"])
        node_off.append(node_off[-1] + len(nodes))
    out = dict(
        extended=extended,
        num_trees=T,
        num_samples=int(meta["numSamples"]),
        num_features=int(meta["numFeatures"]),
        total_num_features=int(meta.get("totalNumFeatures", -1)),
        threshold_score=float(meta["outlierScoreThreshold"]),
        node_off=np.asarray(node_off, np.int32),
        left=np.asarray(left, np.int32),
        right=np.asarray(right, np.int32),
        num_instances=np.asarray(ninst, np.int64),
    )
    if extended:
        out.update(offset=np.asarray(off, np.float64), hp_off=np.asarray(hp_off, np.int64),
                   hp_idx=np.asarray(hp_idx, np.int32), hp_w=np.asarray(hp_w, np.float32))
    else:
        out.update(feature=np.asarray(feat, np.int32), threshold=np.asarray(thr, np.float64))
    return out
