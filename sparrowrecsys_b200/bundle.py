"""TensorFlow-free reader for SavedModel variable bundles.

The reference exports its served model with `tf.keras.models.save_model`
(`TFRecModel/.../NeuralCF.py:97-105`) into
`src/main/resources/webroot/modeldata/<name>/<version>/variables/`:

* `variables.index`  - a LevelDB-format SSTable (uncompressed blocks) mapping
  checkpoint keys to serialized `BundleEntryProto` messages;
* `variables.data-00000-of-00001` - raw little-endian row-major tensors.

Only the subset of both formats that the shipped models use is implemented:
one data shard, no block compression, float32 tensors (other dtypes are listed
but not decoded).  Format facts are restated from SURVEY.md appendix C.
"""
from __future__ import annotations

import os
import struct
from typing import Dict, List, Tuple

import numpy as np

_SSTABLE_MAGIC = 0xDB4775248B80FB57
_DTYPE_FLOAT32 = 1


def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 70:
            raise ValueError("varint too long")


def _read_block(data: bytes, offset: int, size: int) -> bytes:
    block = data[offset:offset + size]
    if len(block) != size:
        raise ValueError("truncated SSTable block")
    ctype = data[offset + size]
    if ctype != 0:
        raise ValueError("compressed SSTable blocks are not supported (type %d)" % ctype)
    return block


def _block_entries(block: bytes) -> List[Tuple[bytes, bytes]]:
    n_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * n_restarts
    pos = 0
    key = b""
    out = []
    while pos < end:
        shared, pos = _varint(block, pos)
        non_shared, pos = _varint(block, pos)
        vlen, pos = _varint(block, pos)
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        out.append((key, block[pos:pos + vlen]))
        pos += vlen
    return out


def _parse_shape(buf: bytes) -> Tuple[int, ...]:
    dims = []
    pos = 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wire = tag >> 3, tag & 7
        if wire == 2:
            ln, pos = _varint(buf, pos)
            sub = buf[pos:pos + ln]
            pos += ln
            if field == 2:                      # Dim
                size = 0
                sp = 0
                while sp < len(sub):
                    t, sp = _varint(sub, sp)
                    if t & 7 == 0:
                        v, sp = _varint(sub, sp)
                        if t >> 3 == 1:
                            size = v
                    elif t & 7 == 2:
                        l2, sp = _varint(sub, sp)
                        sp += l2
                    else:
                        raise ValueError("unexpected wire type in Dim")
                dims.append(size)
        elif wire == 0:
            _, pos = _varint(buf, pos)
        else:
            raise ValueError("unexpected wire type in TensorShapeProto")
    return tuple(dims)


def _parse_entry(buf: bytes) -> dict:
    e = {"dtype": 0, "shape": (), "shard": 0, "offset": 0, "size": 0}
    pos = 0
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wire = tag >> 3, tag & 7
        if wire == 0:
            v, pos = _varint(buf, pos)
            if field == 1:
                e["dtype"] = v
            elif field == 3:
                e["shard"] = v
            elif field == 4:
                e["offset"] = v
            elif field == 5:
                e["size"] = v
        elif wire == 2:
            ln, pos = _varint(buf, pos)
            if field == 2:
                e["shape"] = _parse_shape(buf[pos:pos + ln])
            pos += ln
        elif wire == 5:
            pos += 4
        elif wire == 1:
            pos += 8
        else:
            raise ValueError("unexpected wire type %d" % wire)
    return e


def read_index(index_path: str) -> Dict[str, dict]:
    """Checkpoint key -> {dtype, shape, shard, offset, size}."""
    with open(index_path, "rb") as f:
        data = f.read()
    if len(data) < 48:
        raise ValueError("not an SSTable: too short")
    footer = data[-48:]
    if struct.unpack_from("<Q", footer, 40)[0] != _SSTABLE_MAGIC:
        raise ValueError("not an SSTable: bad magic")
    pos = 0
    _, pos = _varint(footer, pos)        # metaindex handle
    _, pos = _varint(footer, pos)
    idx_off, pos = _varint(footer, pos)
    idx_size, pos = _varint(footer, pos)
    entries: Dict[str, dict] = {}
    for _, handle in _block_entries(_read_block(data, idx_off, idx_size)):
        off, p = _varint(handle, 0)
        size, p = _varint(handle, p)
        for key, value in _block_entries(_read_block(data, off, size)):
            if key == b"":
                continue                 # BundleHeaderProto
            entries[key.decode("utf-8")] = _parse_entry(value)
    return entries


def read_variables(variables_dir: str, include_slots: bool = False) -> Dict[str, np.ndarray]:
    """All float32 model variables of a SavedModel `variables/` directory, keyed by
    the checkpoint key with the `/.ATTRIBUTES/VARIABLE_VALUE` suffix removed and
    the `.S` escape undone.  Optimizer slots and metric accumulators are skipped."""
    index = read_index(os.path.join(variables_dir, "variables.index"))
    data_path = os.path.join(variables_dir, "variables.data-00000-of-00001")
    out: Dict[str, np.ndarray] = {}
    with open(data_path, "rb") as f:
        blob = f.read()
    for key, e in index.items():
        if e["dtype"] != _DTYPE_FLOAT32 or e["shard"] != 0:
            continue
        if not include_slots and (".OPTIMIZER_SLOT" in key or key.startswith("optimizer/")
                                  or key.startswith("keras_api/")):
            continue
        name = key.replace("/.ATTRIBUTES/VARIABLE_VALUE", "").replace(".S", "/")
        n = int(np.prod(e["shape"])) if e["shape"] else 1
        if e["size"] != 4 * n:
            raise ValueError("size mismatch for %s" % key)
        arr = np.frombuffer(blob, dtype="<f4", count=n, offset=e["offset"]).reshape(e["shape"])
        out[name] = np.array(arr, dtype=np.float32)
    return out


def load_neuralcf(savedmodel_dir: str) -> Dict[str, np.ndarray]:
    """Canonical NeuralCF weights (weights.weight_shapes names) from a shipped
    `modeldata/neuralcf/<v>` export (layer order: NeuralCF.py:46-51)."""
    v = read_variables(os.path.join(savedmodel_dir, "variables"))
    g = lambda k: v["layer_with_weights-%s" % k]
    return {
        "movieId_embedding": g("0/movieId_embedding/embedding_weights"),
        "userId_embedding": g("1/userId_embedding/embedding_weights"),
        "dense_0/kernel": g("2/kernel"), "dense_0/bias": g("2/bias"),
        "dense_1/kernel": g("3/kernel"), "dense_1/bias": g("3/bias"),
        "dense_2/kernel": g("4/kernel"), "dense_2/bias": g("4/bias"),
    }


def load_twotowers(savedmodel_dir: str) -> Dict[str, np.ndarray]:
    """Canonical two-tower weights from the shipped `modeldata/MLPRec/005` export:
    one Dense(10, relu) per tower, raw Dot output."""
    v = read_variables(os.path.join(savedmodel_dir, "variables"))
    g = lambda k: v["layer_with_weights-%s" % k]
    return {
        "movieId_embedding": g("0/movieId_embedding/embedding_weights"),
        "userId_embedding": g("1/userId_embedding/embedding_weights"),
        "item_dense_0/kernel": g("2/kernel"), "item_dense_0/bias": g("2/bias"),
        "user_dense_0/kernel": g("3/kernel"), "user_dense_0/bias": g("3/bias"),
    }
