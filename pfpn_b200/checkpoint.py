"""TF-1 tensor-bundle checkpoints (`model.ckpt-N.index` + `.data-0000k-of-0000n`) <-> the flat parameter buffer.

SURVEY section 8f rank 4 / reference `models/utils.py:9-15` (tf.train.Saver): the shipped DPPO-PFPN checkpoints are
tensor bundles whose `.index` is a leveldb table (key = variable name, value = BundleEntryProto{dtype, shape, shard_id,
offset, size, crc32c}; the empty key holds the BundleHeaderProto) and whose `.data-*` shards hold the raw little-endian
tensors.  This module reads and writes that format with numpy only (no TensorFlow), so weights trained with the
reference can be loaded by name into `ParticleFilteringClipPPONetwork` and vice versa.
The reference tree ships only the `.index` files (the `.data` blobs are absent), so the reader is checked against
those indexes (names / shapes / sizes) and the data path by a write -> read round trip.
"""
from __future__ import annotations

import os
import struct
from typing import Dict, NamedTuple, Tuple

import numpy as np

_TABLE_MAGIC = 0xDB4775248B80FB57
_DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64, 10: np.bool_}
_DTYPE_IDS = {np.dtype(v): k for k, v in _DTYPES.items()}


class BundleEntry(NamedTuple):
    dtype: int
    shape: Tuple[int, ...]
    shard_id: int
    offset: int
    size: int
    crc32c: int


# ---- protobuf / varint helpers ------------------------------------------------------------------------------------
def _varint(b: bytes, i: int):
    x = s = 0
    while True:
        c = b[i]
        i += 1
        x |= (c & 0x7F) << s
        if c < 0x80:
            return x, i
        s += 7


def _put_varint(x: int) -> bytes:
    out = bytearray()
    while True:
        c = x & 0x7F
        x >>= 7
        out.append(c | (0x80 if x else 0))
        if not x:
            return bytes(out)


def _fields(b: bytes):
    i = 0
    while i < len(b):
        tag, i = _varint(b, i)
        fno, wt = tag >> 3, tag & 7
        if wt == 0:
            v, i = _varint(b, i)
        elif wt == 1:
            v, i = b[i:i + 8], i + 8
        elif wt == 2:
            n, i = _varint(b, i)
            v, i = b[i:i + n], i + n
        elif wt == 5:
            v, i = b[i:i + 4], i + 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        yield fno, wt, v


def _parse_entry(b: bytes) -> BundleEntry:
    dtype = shard = offset = size = crc = 0
    shape: Tuple[int, ...] = ()
    for fno, wt, v in _fields(b):
        if fno == 1:
            dtype = v
        elif fno == 2:  # TensorShapeProto{ repeated Dim{size=1} dim=2 }
            dims = []
            for f2, _, d in _fields(v):
                if f2 == 2:
                    sz = 0
                    for f3, _, x in _fields(d):
                        if f3 == 1:
                            sz = x
                    dims.append(sz)
            shape = tuple(dims)
        elif fno == 3:
            shard = v
        elif fno == 4:
            offset = v
        elif fno == 5:
            size = v
        elif fno == 6:
            crc = struct.unpack("<I", v)[0]
    return BundleEntry(dtype, shape, shard, offset, size, crc)


def _encode_entry(e: BundleEntry) -> bytes:
    out = bytearray()
    out += b"\x08" + _put_varint(e.dtype)
    dims = b"".join(b"\x12" + _put_varint(len(d)) + d for d in (b"\x08" + _put_varint(s) for s in e.shape))
    out += b"\x12" + _put_varint(len(dims)) + dims
    if e.shard_id:
        out += b"\x18" + _put_varint(e.shard_id)
    if e.offset:
        out += b"\x20" + _put_varint(e.offset)
    out += b"\x28" + _put_varint(e.size)
    out += b"\x35" + struct.pack("<I", e.crc32c)
    return bytes(out)


# ---- crc32c (Castagnoli), masked as leveldb / TF store it ---------------------------------------------------------
def _make_crc_tables():
    """Slice-by-8 tables: T[0] is the byte-wise table, T[k][n] = crc of byte n followed by k zero bytes."""
    t0 = []
    for n in range(256):
        c = n
        for _ in range(8):
            c = (c >> 1) ^ 0x82F63B78 if c & 1 else c >> 1
        t0.append(c)
    tabs = [t0]
    for _ in range(7):
        prev = tabs[-1]
        tabs.append([(prev[n] >> 8) ^ t0[prev[n] & 0xFF] for n in range(256)])
    return tabs


_CRC_TABLES = _make_crc_tables()


def crc32c(data: bytes) -> int:
    """Castagnoli CRC, slice-by-8 (one Python iteration per 8 bytes; ~6 MB bundles in well under a second)."""
    t0, t1, t2, t3, t4, t5, t6, t7 = _CRC_TABLES
    c = 0xFFFFFFFF
    n8 = len(data) & ~7
    for (q,) in struct.iter_unpack("<Q", memoryview(data)[:n8]):
        q ^= c
        c = (t7[q & 0xFF] ^ t6[(q >> 8) & 0xFF] ^ t5[(q >> 16) & 0xFF] ^ t4[(q >> 24) & 0xFF] ^
             t3[(q >> 32) & 0xFF] ^ t2[(q >> 40) & 0xFF] ^ t1[(q >> 48) & 0xFF] ^ t0[q >> 56])
    for byte in memoryview(data)[n8:]:
        c = t0[(c ^ byte) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def _mask(crc: int) -> int:
    return ((((crc >> 15) | (crc << 17)) & 0xFFFFFFFF) + 0xA282EAD8) & 0xFFFFFFFF


# ---- leveldb table ------------------------------------------------------------------------------------------------
def _read_block(b: bytes, off: int, size: int):
    blk = b[off:off + size]
    if b[off + size] != 0:
        raise ValueError("compressed index blocks are not supported (TF writes tensor-bundle indexes uncompressed)")
    nrestart = struct.unpack("<I", blk[-4:])[0]
    end = len(blk) - 4 - 4 * nrestart
    i, key = 0, b""
    while i < end:
        shared, i = _varint(blk, i)
        non_shared, i = _varint(blk, i)
        vlen, i = _varint(blk, i)
        key = key[:shared] + blk[i:i + non_shared]
        i += non_shared
        yield key, blk[i:i + vlen]
        i += vlen


def read_bundle_index(index_path: str) -> Dict[str, BundleEntry]:
    """Variable name -> BundleEntry for every tensor of a `.index` file (header entry excluded)."""
    b = open(index_path, "rb").read()
    if len(b) < 48 or struct.unpack("<Q", b[-8:])[0] != _TABLE_MAGIC:
        raise ValueError(f"{index_path}: not a leveldb table / tensor-bundle index")
    footer = b[-48:]
    _, i = _varint(footer, 0)          # metaindex handle: offset
    _, i = _varint(footer, i)          #                   size
    ioff, i = _varint(footer, i)       # index handle
    isize, i = _varint(footer, i)
    out: Dict[str, BundleEntry] = {}
    for _, handle in _read_block(b, ioff, isize):
        doff, j = _varint(handle, 0)
        dsize, j = _varint(handle, j)
        for key, val in _read_block(b, doff, dsize):
            if key:
                out[key.decode()] = _parse_entry(val)
    return out


def load_bundle(prefix: str) -> Dict[str, np.ndarray]:
    """All tensors of the checkpoint `prefix` (e.g. '.../model.ckpt-78000') as numpy arrays."""
    entries = read_bundle_index(prefix + ".index")
    nshards = 1 + max((e.shard_id for e in entries.values()), default=0)
    shards = {}
    out = {}
    for name, e in entries.items():
        if e.shard_id not in shards:
            shards[e.shard_id] = open(f"{prefix}.data-{e.shard_id:05d}-of-{nshards:05d}", "rb").read()
        raw = shards[e.shard_id][e.offset:e.offset + e.size]
        if len(raw) != e.size:
            raise ValueError(f"{name}: data shard truncated")
        if e.crc32c and _mask(crc32c(raw)) != e.crc32c:
            raise ValueError(f"{name}: crc32c mismatch")
        out[name] = np.frombuffer(raw, dtype=_DTYPES[e.dtype]).reshape(e.shape).copy()
    return out


def _block(entries) -> bytes:
    """One uncompressed block, restart interval 1 (no key prefix sharing), + trailer."""
    body, restarts = bytearray(), []
    for key, val in entries:
        restarts.append(len(body))
        body += _put_varint(0) + _put_varint(len(key)) + _put_varint(len(val)) + key + val
    if not restarts:
        restarts = [0]
    for r in restarts:
        body += struct.pack("<I", r)
    body += struct.pack("<I", len(restarts))
    trailer = b"\x00" + struct.pack("<I", _mask(crc32c(bytes(body) + b"\x00")))
    return bytes(body) + trailer


def write_bundle(prefix: str, tensors: Dict[str, np.ndarray]) -> None:
    """Write `tensors` as a single-shard TF-1 tensor bundle (`prefix`.index + .data-00000-of-00001)."""
    os.makedirs(os.path.dirname(os.path.abspath(prefix)), exist_ok=True)
    data, items = bytearray(), []
    for name in sorted(tensors):
        a = np.asarray(tensors[name])  # (ascontiguousarray would promote scalars to shape (1,))
        if not a.flags.c_contiguous:
            a = a.copy(order="C")
        raw = a.astype(a.dtype.newbyteorder("<"), copy=False).tobytes()
        e = BundleEntry(_DTYPE_IDS[np.dtype(a.dtype)], tuple(int(s) for s in a.shape), 0, len(data), len(raw), _mask(crc32c(raw)))
        items.append((name.encode(), _encode_entry(e)))
        data += raw
    header = b"\x08\x01\x1a\x02\x08\x01"  # BundleHeaderProto{num_shards: 1, version{producer: 1}} (as the shipped files)
    blk = _block([(b"", header)] + items)
    meta = _block([])
    last_key = items[-1][0] if items else b""
    index = _block([(last_key, _put_varint(0) + _put_varint(len(blk) - 5))])
    footer = _put_varint(len(blk)) + _put_varint(len(meta) - 5) + _put_varint(len(blk) + len(meta)) + _put_varint(len(index) - 5)
    footer = footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", _TABLE_MAGIC)
    with open(prefix + ".index", "wb") as f:
        f.write(blk + meta + index + footer)
    with open(prefix + ".data-00000-of-00001", "wb") as f:
        f.write(bytes(data))


# ---- the network ---------------------------------------------------------------------------------------------------
_STATE_NAMES = {  # reference variable name -> attribute of ParticleFilteringClipPPONetwork (names from the shipped .index)
    "global_net/max_active_degree": "max_active",
    "global_net/sum_active_degree": "sum_active",
    "global_net/state_normalizer/mean": "state_mean",
    "global_net/state_normalizer/std": "state_std",
}


def network_tensors(net) -> Dict[str, np.ndarray]:
    """The reference's variable names -> host arrays for everything `net` owns (weights, particles, statistics,
    resample counter, global step)."""
    out = {k: p.detach().cpu().numpy().copy() for k, (p, _) in net.named_parameters().items()}
    for name, attr in _STATE_NAMES.items():
        out[name] = getattr(net, attr).detach().cpu().numpy().copy()
    out["global_net/resample/train_flag"] = np.float32(net.train_flag)      # a float32 scalar in the reference graph
    out["step/global_step"] = np.int64(net.global_step)
    return out


def save_tf_checkpoint(net, prefix: str) -> None:
    write_bundle(prefix, {k: np.asarray(v) for k, v in network_tensors(net).items()})


def load_tf_checkpoint(net, prefix: str, strict: bool = True):
    """Copy every variable of the bundle that `net` knows (by reference name) into the network.
    Returns (loaded names, bundle names that the network does not own, e.g. Adam slots / beta powers / worker copies)."""
    import torch
    tensors = load_bundle(prefix)
    params = net.named_parameters()
    loaded, extra = [], []
    for name, arr in tensors.items():
        if name in params:
            dst = params[name][0]
        elif name in _STATE_NAMES:
            dst = getattr(net, _STATE_NAMES[name])
        elif name == "global_net/resample/train_flag":
            net.train_flag = int(arr)
            loaded.append(name)
            continue
        elif name == "step/global_step":
            net.global_step = int(arr)
            loaded.append(name)
            continue
        else:
            extra.append(name)
            continue
        if tuple(dst.shape) != tuple(arr.shape):
            raise ValueError(f"{name}: checkpoint shape {arr.shape} != network shape {tuple(dst.shape)}")
        dst.copy_(torch.from_numpy(np.ascontiguousarray(arr)).to(dst.device))
        loaded.append(name)
    missing = [k for k in params if k not in tensors]
    if strict and missing:
        raise KeyError(f"checkpoint lacks {missing[:5]}{'...' if len(missing) > 5 else ''}")
    return loaded, extra
