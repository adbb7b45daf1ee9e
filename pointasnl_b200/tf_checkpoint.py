"""TensorFlow V2 checkpoint ("tensor bundle") import / export without TensorFlow.

The reference saves and restores with `tf.train.Saver` (train.py:157,170-172,198-200; released weights: README.md:94):
`<prefix>.index` + `<prefix>.data-00000-of-00001`.  The variable names this package creates are the reference's own
(`layer1/conv0/weights`, `layer1/conv0/bn/moving_mean`, `fc1/biases`, ...; tf_util.VariableStore) with the same
shapes, so a checkpoint maps onto a VariableStore by NAME, no table needed; optimizer slots (`.../Adam`, `.../Adam_1`,
`beta1_power`, ...) and the global step are ignored on import.

Format (tensorflow/core/util/tensor_bundle, tensorflow/core/lib/io/table -- restated from the published format, TF is
not installable here, so this reader is checked by round trip with its own writer and against hand-built index
blocks, not against a file TF wrote: PARITY UNPINNED):
  * `.index` is a LevelDB-style sorted table: blocks of prefix-compressed (key, value) entries
    [shared varint32 | non_shared varint32 | value_len varint32 | key suffix | value] followed by a uint32 restart array
    and its length, each block trailed by 1 compression byte (0 = none, 1 = snappy) + masked crc32c; a 48-byte footer
    holds the metaindex and index block handles (varint64 offset, size) and the magic 0xdb4775248b80fb57.
  * key "" -> BundleHeaderProto {1: num_shards, 2: endianness, 3: version}; every other key is a tensor name ->
    BundleEntryProto {1: dtype, 2: TensorShapeProto{2: dim{1: size}}, 3: shard_id, 4: offset, 5: size, 6: crc32c}.
  * `.data-SSSSS-of-NNNNN` holds the raw little-endian tensor bytes at [offset, offset+size).
"""
import os
import struct

import numpy as np

TABLE_MAGIC = 0xDB4775248B80FB57
DT_FLOAT, DT_DOUBLE, DT_INT32, DT_INT64 = 1, 2, 3, 9
_NP_OF = {DT_FLOAT: np.float32, DT_DOUBLE: np.float64, DT_INT32: np.int32, DT_INT64: np.int64}
_DT_OF = {np.dtype(np.float32): DT_FLOAT, np.dtype(np.float64): DT_DOUBLE, np.dtype(np.int32): DT_INT32,
          np.dtype(np.int64): DT_INT64}


# ------------------------------------------------------------------------------------------------ varints / protobuf
def _read_varint(buf, pos):
    shift = result = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _parse_message(buf):
    """Minimal protobuf wire parser: field number -> list of values (ints for varint / fixed, bytes for length-delimited)."""
    fields, pos = {}, 0
    while pos < len(buf):
        tag, pos = _read_varint(buf, pos)
        num, wt = tag >> 3, tag & 7
        if wt == 0:
            val, pos = _read_varint(buf, pos)
        elif wt == 1:
            val = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        elif wt == 2:
            ln, pos = _read_varint(buf, pos)
            val = bytes(buf[pos:pos + ln])
            pos += ln
        elif wt == 5:
            val = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        else:
            raise ValueError(f"unsupported protobuf wire type {wt}")
        fields.setdefault(num, []).append(val)
    return fields


def _signed64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _parse_entry(buf):
    f = _parse_message(buf)
    shape = []
    if 2 in f:
        for dim in _parse_message(f[2][0]).get(2, []):
            shape.append(_signed64(_parse_message(dim).get(1, [0])[0]))
    return {"dtype": f.get(1, [0])[0], "shape": tuple(shape), "shard_id": f.get(3, [0])[0], "offset": f.get(4, [0])[0],
            "size": f.get(5, [0])[0], "crc32c": f.get(6, [None])[0], "sliced": 7 in f}


def _field(num, wt, payload):
    return _varint((num << 3) | wt) + payload


def _encode_entry(dtype, shape, offset, size, crc):
    dims = b"".join(_field(2, 2, _varint(len(d)) + d) for d in (_field(1, 0, _varint(int(s))) for s in shape))
    msg = _field(1, 0, _varint(dtype)) + _field(2, 2, _varint(len(dims)) + dims)
    if offset:
        msg += _field(4, 0, _varint(offset))
    msg += _field(5, 0, _varint(size)) + _field(6, 5, struct.pack("<I", crc))
    return msg


# ------------------------------------------------------------------------------------------------ crc32c (Castagnoli)
_CRC_TABLE = None


def crc32c(data, crc=0):
    global _CRC_TABLE
    if _CRC_TABLE is None:
        tbl = []
        for i in range(256):
            c = i
            for _ in range(8):
                c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
            tbl.append(c)
        _CRC_TABLE = np.array(tbl, dtype=np.uint32)
    c = crc ^ 0xFFFFFFFF
    tbl = _CRC_TABLE
    for b in bytes(data):
        c = int(tbl[(c ^ b) & 0xFF]) ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def _mask_crc(c):
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


# ------------------------------------------------------------------------------------------------ table (.index) reader
def _block_entries(block):
    """(key, value) pairs of one uncompressed table block."""
    num_restarts = struct.unpack_from("<I", block, len(block) - 4)[0]
    limit = len(block) - 4 - 4 * num_restarts
    pos, key, out = 0, b"", []
    while pos < limit:
        shared, pos = _read_varint(block, pos)
        non_shared, pos = _read_varint(block, pos)
        vlen, pos = _read_varint(block, pos)
        key = key[:shared] + bytes(block[pos:pos + non_shared])
        pos += non_shared
        out.append((key, bytes(block[pos:pos + vlen])))
        pos += vlen
    return out


def _read_block(data, offset, size):
    ctype = data[offset + size]
    if ctype == 1:
        raise NotImplementedError("snappy-compressed table blocks (tf.train.Saver writes checkpoints uncompressed)")
    if ctype != 0:
        raise ValueError(f"unknown block compression type {ctype}")
    return data[offset:offset + size]


def read_index(path):
    """`<prefix>.index` -> (header dict, {tensor name: entry dict})."""
    data = open(path, "rb").read()
    if len(data) < 48 or struct.unpack_from("<Q", data, len(data) - 8)[0] != TABLE_MAGIC:
        raise ValueError(f"{path}: not a TensorFlow checkpoint index (bad table magic)")
    footer = data[-48:]
    _mo, pos = _read_varint(footer, 0)
    _ms, pos = _read_varint(footer, pos)
    io, pos = _read_varint(footer, pos)
    isz, pos = _read_varint(footer, pos)
    entries, header = {}, {}
    for _sep, handle in _block_entries(_read_block(data, io, isz)):
        off, p = _read_varint(handle, 0)
        size, p = _read_varint(handle, p)
        for key, value in _block_entries(_read_block(data, off, size)):
            if key == b"":
                h = _parse_message(value)
                header = {"num_shards": h.get(1, [1])[0], "endianness": h.get(2, [0])[0]}
            else:
                entries[key.decode()] = _parse_entry(value)
    return header, entries


def load_checkpoint(prefix, names=None, skip_optimizer_slots=True):
    """`prefix` as passed to saver.save / saver.restore (e.g. 'log/best_model.ckpt') -> {name: np.ndarray}."""
    header, entries = read_index(prefix + ".index")
    if header.get("endianness", 0) != 0:
        raise NotImplementedError("big-endian checkpoint")
    nshards = header.get("num_shards", 1)
    shards = {}
    out = {}
    for name, e in entries.items():
        if names is not None and name not in names:
            continue
        if skip_optimizer_slots and (name.endswith("/Adam") or name.endswith("/Adam_1") or name.endswith("/Momentum") or
                                     name in ("beta1_power", "beta2_power", "Variable", "global_step")):
            continue
        if e["sliced"]:
            raise NotImplementedError(f"{name}: partitioned (sliced) variable")
        if e["dtype"] not in _NP_OF:
            continue
        sid = e["shard_id"]
        if sid not in shards:
            shards[sid] = np.memmap(f"{prefix}.data-{sid:05d}-of-{nshards:05d}", dtype=np.uint8, mode="r")
        raw = np.asarray(shards[sid][e["offset"]:e["offset"] + e["size"]])
        arr = raw.view(_NP_OF[e["dtype"]]).reshape(e["shape"]).copy()
        out[name] = arr
    return out


# ------------------------------------------------------------------------------------------------ writer
def _build_block(pairs, restart_interval=16):
    out, restarts, prev = bytearray(), [], b""
    for i, (key, value) in enumerate(pairs):
        if i % restart_interval == 0:
            restarts.append(len(out))
            shared = 0
        else:
            shared = 0
            while shared < min(len(prev), len(key)) and prev[shared] == key[shared]:
                shared += 1
        out += _varint(shared) + _varint(len(key) - shared) + _varint(len(value)) + key[shared:] + value
        prev = key
    if not restarts:
        restarts = [0]
    for r in restarts:
        out += struct.pack("<I", r)
    out += struct.pack("<I", len(restarts))
    return bytes(out)


def _emit_block(f, block):
    off = f.tell()
    f.write(block)
    f.write(b"\x00")
    f.write(struct.pack("<I", _mask_crc(crc32c(block + b"\x00"))))
    return off, len(block)


def save_checkpoint(prefix, variables, entries_per_block=64):
    """{name: array-like} -> `<prefix>.index` + `<prefix>.data-00000-of-00001` (one shard, uncompressed), the layout
    tf.train.Saver(write_version=V2) produces; usable by the reference's saver.restore."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    names = sorted(variables)
    offset, meta = 0, {}
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        for n in names:
            a = np.ascontiguousarray(np.asarray(variables[n]))
            if a.dtype not in _DT_OF:
                a = a.astype(np.float32)
            raw = a.tobytes()
            f.write(raw)
            meta[n] = (_DT_OF[a.dtype], a.shape, offset, len(raw), _mask_crc(crc32c(raw)))
            offset += len(raw)
    header = _field(1, 0, _varint(1)) + _field(3, 2, (lambda v: _varint(len(v)) + v)(_field(1, 0, _varint(1))))
    pairs = [(b"", header)] + [(n.encode(), _encode_entry(*meta[n])) for n in names]
    with open(prefix + ".index", "wb") as f:
        index_pairs = []
        for i in range(0, len(pairs), entries_per_block):
            chunk = pairs[i:i + entries_per_block]
            off, size = _emit_block(f, _build_block(chunk))
            index_pairs.append((chunk[-1][0] + b"\x00" if i + entries_per_block < len(pairs) else chunk[-1][0] + b"\xff",
                                _varint(off) + _varint(size)))
        moff, msize = _emit_block(f, _build_block([]))
        ioff, isize = _emit_block(f, _build_block(index_pairs, restart_interval=1))
        footer = _varint(moff) + _varint(msize) + _varint(ioff) + _varint(isize)
        f.write(footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC))


# ------------------------------------------------------------------------------------------------ VariableStore glue
def restore(store, prefix, strict=True):
    """Load a checkpoint into an already materialised tf_util.VariableStore (same names, same shapes).  Returns the
    list of checkpoint tensors that found no variable.  strict: every variable of the store must be in the file."""
    values = load_checkpoint(prefix)
    have = store.variables()
    missing = [n for n in have if n not in values]
    if strict and missing:
        raise KeyError(f"checkpoint {prefix} lacks {len(missing)} variables, e.g. {missing[:3]}")
    for n, v in have.items():
        if n in values and tuple(values[n].shape) != tuple(v.shape):
            raise ValueError(f"{n}: checkpoint shape {values[n].shape} != variable shape {tuple(v.shape)}")
    store.load_variables({n: values[n] for n in have if n in values})
    return [n for n in values if n not in have]


def initializer_from_checkpoint(prefix):
    """init_fn for tf_util.VariableStore(init_fn=...): variables take their checkpoint value at creation time."""
    values = load_checkpoint(prefix)

    def init_fn(full_name, shape, leaf):
        v = values.get(full_name)
        if v is not None and tuple(v.shape) != tuple(shape):
            raise ValueError(f"{full_name}: checkpoint shape {v.shape} != requested {tuple(shape)}")
        return v
    return init_fn


def save(store, prefix):
    save_checkpoint(prefix, {n: v.detach().cpu().numpy() for n, v in store.variables().items()})
