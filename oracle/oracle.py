"""CPU oracle for the model-hub weight-load path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module.  The product (``kukeon_b200`` / ``libkukeon_gpuload.so``) never does, and has
no CPU fallback.

PARITY UNPINNED AGAINST THE REFERENCE: eminwux/kukeon @ 4be245a contains no weight loader at all
(``internal/modelhub`` is the orchestrator's resource data model — ``internal/modelhub/cell.go:21``,
``container.go:21``) and its ``go.mod`` pulls no safetensors/GGUF module, so there is no
reference code, call site, test or golden vector for this path.  The oracle is therefore a restatement of
the *published file formats*, pinned instead against the format owners' own libraries installed in the
image:

* safetensors 0.7.0  — header layout, validation rules  (``tests/test_index.py``)
* gguf 0.19.0        — GGUF v3 header, ``quants.{Q4_0,Q4_1,Q5_0,Q5_1,Q8_0,Q2_K,Q3_K,Q4_K,Q5_K,Q6_K}.dequantize_blocks`` (gguf/quants.py:220-572)
* torch 2.11         — fp32/fp16 -> bf16 round-to-nearest-even

Everything here is plain Python/numpy so that it can be read next to those sources.  Heavy loops have a
C twin in ``oracle/kk_oracle.c`` (same arithmetic, OpenMP) used for full-size checks and the CPU baseline.
"""
from __future__ import annotations

import json
import os
import struct
from typing import Dict, List, Tuple

import numpy as np

# ---------------------------------------------------------------------------------------------
# dtypes
# ---------------------------------------------------------------------------------------------
# safetensors Dtype names -> bits per element (safetensors/src/tensor.rs `Dtype::bitsize`)
ST_BITS = {
    "BOOL": 8, "F4": 4, "F6_E2M3": 6, "F6_E3M2": 6, "U8": 8, "I8": 8, "F8_E5M2": 8, "F8_E4M3": 8,
    "F8_E8M0": 8, "I16": 16, "U16": 16, "F16": 16, "BF16": 16, "I32": 32, "U32": 32, "F32": 32,
    "C64": 64, "F64": 64, "I64": 64, "U64": 64,
}
# ggml type id -> (name, elements per block, bytes per block)   (gguf/constants.py GGML_QUANT_SIZES)
GGML_TYPES = {
    0: ("F32", 1, 4), 1: ("F16", 1, 2), 2: ("Q4_0", 32, 18), 3: ("Q4_1", 32, 20), 6: ("Q5_0", 32, 22),
    7: ("Q5_1", 32, 24), 8: ("Q8_0", 32, 34), 10: ("Q2_K", 256, 84), 11: ("Q3_K", 256, 110),
    12: ("Q4_K", 256, 144), 13: ("Q5_K", 256, 176), 14: ("Q6_K", 256, 210), 15: ("Q8_K", 256, 292),
    20: ("IQ4_NL", 32, 18), 23: ("IQ4_XS", 256, 136), 39: ("MXFP4", 32, 17), 16: ("IQ2_XXS", 256, 66), 17: ("IQ2_XS", 256, 74),
    18: ("IQ3_XXS", 256, 98), 19: ("IQ1_S", 256, 50), 21: ("IQ3_S", 256, 110), 22: ("IQ2_S", 256, 82), 29: ("IQ1_M", 256, 56),
    34: ("TQ1_0", 256, 54), 35: ("TQ2_0", 256, 66), 40: ("NVFP4", 64, 36),
    24: ("I8", 1, 1), 25: ("I16", 1, 2), 26: ("I32", 1, 4), 27: ("I64", 1, 8), 28: ("F64", 1, 8),
    30: ("BF16", 1, 2),
}
POOL_ALIGN = 256


class OracleError(ValueError):
    """Malformed checkpoint (what the product reports as KK_EFORMAT / KK_EUNSUPPORTED)."""


# ---------------------------------------------------------------------------------------------
# index ("Pull")
# ---------------------------------------------------------------------------------------------
def index_safetensors(path: str, shard: int = 0) -> List[dict]:
    """safetensors layout: u64 LE N | N bytes JSON | data.  Validation = the reader's rules probed in
    SURVEY.md Appendix C.3: offsets sorted, gap-free from 0, covering the data section exactly;
    byte size == prod(shape) * dtype bits / 8; header <= 100,000,000 bytes."""
    size = os.path.getsize(path)
    with open(path, "rb") as f:
        head = f.read(8)
        if len(head) < 8:
            raise OracleError("header too small")
        (n,) = struct.unpack("<Q", head)
        if n > 100_000_000:
            raise OracleError("header too large")
        if n > size - 8:
            raise OracleError("invalid header length")
        raw = f.read(n)
    if not raw.startswith(b"{"):
        raise OracleError("invalid header start")
    try:
        pairs = json.loads(raw.decode("utf-8"), object_pairs_hook=list)
    except Exception as e:  # noqa: BLE001
        raise OracleError(f"invalid header deserialization: {e}") from e
    ents = []
    seen = set()
    for name, info in pairs:
        if name == "__metadata__":
            if not isinstance(info, list) or any(not isinstance(v, str) for _, v in info):
                raise OracleError("bad __metadata__")
            continue
        if name in seen:
            raise OracleError(f"duplicate tensor {name}")
        seen.add(name)
        d = dict(info) if isinstance(info, list) else None
        if d is None or "dtype" not in d or "shape" not in d or "data_offsets" not in d:
            raise OracleError(f"bad tensor entry {name}")
        if d["dtype"] not in ST_BITS:
            raise OracleError(f"unknown variant `{d['dtype']}`")
        b, e = d["data_offsets"]
        ents.append((b, e, name, d["dtype"], [int(x) for x in d["shape"]]))
    ents.sort(key=lambda t: (t[0], t[1]))
    data_start = 8 + n
    cur = 0
    out = []
    for b, e, name, dt, shape in ents:
        if b != cur or e < b:
            raise OracleError(f"invalid offset for tensor `{name}`")
        cur = e
        nel = 1
        for s in shape:
            nel *= s
        nbits = nel * ST_BITS[dt]
        if nbits % 8 or nbits // 8 != e - b:
            raise OracleError(f"invalid shape, data type, or offset for tensor `{name}`")
        out.append(dict(name=name, dtype=dt, shape=shape, shard=shard, file_offset=data_start + b, nbytes=e - b))
    if data_start + cur != size:
        raise OracleError("incomplete metadata, file not fully covered")
    return out


def _gguf_skip(buf: memoryview, off: int, vt: int) -> int:
    scalar = {0: 1, 1: 1, 7: 1, 2: 2, 3: 2, 4: 4, 5: 4, 6: 4, 10: 8, 11: 8, 12: 8}
    if vt in scalar:
        return off + scalar[vt]
    if vt == 8:
        (ln,) = struct.unpack_from("<Q", buf, off)
        return off + 8 + ln
    if vt == 9:
        et, cnt = struct.unpack_from("<IQ", buf, off)
        off += 12
        if et in scalar:
            return off + scalar[et] * cnt
        for _ in range(cnt):
            off = _gguf_skip(buf, off, et)
        return off
    raise OracleError(f"unknown GGUF value type {vt}")


def index_gguf(path: str, shard: int = 0) -> List[dict]:
    """GGUF v2/v3 (gguf-py gguf_reader.py:132-186, :259-345): `<IIQQ` magic/version/n_tensors/n_kv, KVs,
    tensor infos (name, n_dims, ne[] innermost-first, ggml type, offset), data at the next multiple of
    general.alignment (default 32).  Shapes are reported outermost-first."""
    size = os.path.getsize(path)
    with open(path, "rb") as f:
        buf = memoryview(f.read(min(size, 1 << 28)))  # headers are far smaller than 256 MiB
    if len(buf) < 24:
        raise OracleError("too small for a GGUF header")
    magic, version, n_tensors, n_kv = struct.unpack_from("<IIQQ", buf, 0)
    if magic != 0x46554747:
        raise OracleError("GGUF magic invalid")
    if version not in (2, 3):
        raise OracleError(f"GGUF version {version} unsupported")
    off = 24
    alignment = 32
    try:
        for _ in range(n_kv):
            (kl,) = struct.unpack_from("<Q", buf, off)
            key = bytes(buf[off + 8: off + 8 + kl]).decode("utf-8")
            off += 8 + kl
            (vt,) = struct.unpack_from("<I", buf, off)
            off += 4
            if key == "general.alignment":
                if vt != 4:
                    raise OracleError("bad type for general.alignment")
                (alignment,) = struct.unpack_from("<I", buf, off)
                if alignment == 0 or alignment & (alignment - 1):
                    raise OracleError("invalid alignment")
            off = _gguf_skip(buf, off, vt)
        infos = []
        for _ in range(n_tensors):
            (nl,) = struct.unpack_from("<Q", buf, off)
            name = bytes(buf[off + 8: off + 8 + nl]).decode("utf-8")
            off += 8 + nl
            (nd,) = struct.unpack_from("<I", buf, off)
            off += 4
            ne = list(struct.unpack_from(f"<{nd}Q", buf, off))
            off += 8 * nd
            gt, rel = struct.unpack_from("<IQ", buf, off)
            off += 12
            infos.append((name, ne, gt, rel))
    except struct.error as e:
        raise OracleError(f"truncated GGUF header: {e}") from e
    data_start = (off + alignment - 1) // alignment * alignment
    out = []
    names = set()
    for name, ne, gt, rel in infos:
        if name in names:
            raise OracleError(f"duplicated tensor {name}")
        names.add(name)
        if gt not in GGML_TYPES:
            raise OracleError(f"ggml type {gt} unsupported")
        tname, bel, bby = GGML_TYPES[gt]
        nel = 1
        for d in ne:
            nel *= d
        if bel > 1 and (not ne or ne[0] % bel):
            raise OracleError(f"tensor {name}: row not a multiple of the block size")
        nbytes = nel // bel * bby
        if rel % alignment:
            raise OracleError(f"tensor {name}: misaligned data offset")
        if data_start + rel + nbytes > size:
            raise OracleError(f"tensor {name}: data out of file bounds")
        out.append(dict(name=name, dtype=tname, shape=list(reversed(ne)), shard=shard,
                        file_offset=data_start + rel, nbytes=nbytes))
    return out


def _canon_sort(recs: List[dict]) -> List[dict]:
    return sorted(recs, key=lambda r: (r["shard"], r["file_offset"], r["nbytes"], r["name"]))


def index_path(path: str) -> Tuple[List[str], List[dict]]:
    """Resolve a checkpoint path exactly the way kk_index documents it; returns (shard paths, records
    sorted by (shard, file_offset))."""
    path = os.path.realpath(path)
    if os.path.isdir(path):
        idx = os.path.join(path, "model.safetensors.index.json")
        if os.path.isfile(idx):
            return _index_sharded(idx)
        if os.path.isfile(os.path.join(path, "model.safetensors")):
            files, kind = ["model.safetensors"], "st"
        else:
            st = sorted(f for f in os.listdir(path) if f.endswith(".safetensors") and os.path.isfile(os.path.join(path, f)))
            gg = sorted(f for f in os.listdir(path) if f.endswith(".gguf") and os.path.isfile(os.path.join(path, f)))
            if st:
                files, kind = st, "st"
            elif gg:
                files, kind = gg, "gguf"
            else:
                raise FileNotFoundError(path)
        shards = [os.path.join(path, f) for f in files]
    else:
        if not os.path.isfile(path):
            raise FileNotFoundError(path)
        if path.endswith(".index.json"):
            return _index_sharded(path)
        with open(path, "rb") as f:
            sniff = f.read(4)
        kind = "gguf" if (path.endswith(".gguf") or sniff == b"GGUF") and not path.endswith(".safetensors") else "st"
        shards = [path]
    recs: List[dict] = []
    for i, s in enumerate(shards):
        recs += index_gguf(s, i) if kind == "gguf" else index_safetensors(s, i)
    _check_unique(recs)
    return shards, _canon_sort(recs)


def _check_unique(recs):
    names = set()
    for r in recs:
        if r["name"] in names:
            raise OracleError(f"tensor {r['name']} appears in more than one shard")
        names.add(r["name"])


def _index_sharded(index_json: str) -> Tuple[List[str], List[dict]]:
    """HF sharded layout: weight_map name -> shard file; shards numbered in
    file-name order."""
    with open(index_json, "rb") as f:
        doc = json.loads(f.read().decode("utf-8"))
    wm = doc.get("weight_map") if isinstance(doc, dict) else None
    if not isinstance(wm, dict) or not wm:
        raise OracleError("no weight_map")
    files = sorted(set(wm.values()))
    d = os.path.dirname(index_json)
    shards = [os.path.join(d, f) for f in files]
    recs: List[dict] = []
    for i, s in enumerate(shards):
        recs += index_safetensors(s, i)
    _check_unique(recs)
    where = {r["name"]: r["shard"] for r in recs}
    for name, fn in wm.items():
        if name not in where or where[name] != files.index(fn):
            raise OracleError(f"weight_map entry {name} does not match shard contents")
    return shards, _canon_sort(recs)


# ---------------------------------------------------------------------------------------------
# value conversions (bit-level, numpy)
# ---------------------------------------------------------------------------------------------
def f32_bits_to_bf16(u32: np.ndarray) -> np.ndarray:
    """fp32 bit patterns -> bf16 bit patterns, round-to-nearest-even; any NaN -> 0x7FFF (the canonical
    NaN of PTX `cvt.rn.bf16x2.f32`, which the product uses; torch's CPU cast agrees on every non-NaN)."""
    u = u32.astype(np.uint32, copy=False)
    nan = (u & np.uint32(0x7FFFFFFF)) > np.uint32(0x7F800000)
    lsb = (u >> np.uint32(16)) & np.uint32(1)
    r = ((u.astype(np.uint64) + np.uint64(0x7FFF) + lsb.astype(np.uint64)) >> np.uint64(16)).astype(np.uint16)
    r[nan] = 0x7FFF
    return r


def f32_to_bf16(x: np.ndarray) -> np.ndarray:
    return f32_bits_to_bf16(np.ascontiguousarray(x, dtype=np.float32).view(np.uint32))


def f16_bits_to_bf16(u16: np.ndarray) -> np.ndarray:
    """fp16 -> fp32 is exact (subnormals become normals), then RNE to bf16."""
    return f32_to_bf16(np.ascontiguousarray(u16, dtype=np.uint16).view(np.float16).astype(np.float32))


def f8e4m3_bits_to_bf16(u8: np.ndarray) -> np.ndarray:
    """OCP FP8 E4M3 ("fn": no infinities, S.1111.111 is NaN) -> bf16.  Exact: 3 mantissa bits and exponents 2^-9..2^8 all fit.
    Normal: sign | (e + 120) << 7 | m << 4 (bias 7 -> 127); subnormal (e = 0): m * 2^-9; NaN -> 0x7FFF like every cast here."""
    u = np.ascontiguousarray(u8, dtype=np.uint8).astype(np.uint16)
    sgn, e, m = (u >> 7) << 15, (u >> 3) & 15, u & 7
    out = sgn | ((e + 120) << 7) | (m << 4)
    sub = e == 0
    out[sub] = sgn[sub] | f32_to_bf16(m[sub].astype(np.float32) * np.float32(2.0 ** -9))
    out[(e == 15) & (m == 7)] = 0x7FFF
    return out.astype(np.uint16)


def f8e5m2_bits_to_bf16(u8: np.ndarray) -> np.ndarray:
    """FP8 E5M2 is the top byte of an IEEE half: widen through fp16 (exact; +-inf kept, NaN -> 0x7FFF)."""
    return f16_bits_to_bf16(np.ascontiguousarray(u8, dtype=np.uint8).astype(np.uint16) << 8)


def q4k_scale_min(scales: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """6-bit packed (scale, min) of the 8 sub-blocks; gguf/quants.py:479-501 restated with explicit
    indices (j<4: sc=s[j]&63, m=s[j+4]&63; j>=4: sc=(s[j+4]&15)|((s[j-4]>>6)<<4), m=(s[j+4]>>4)|((s[j]>>6)<<4))."""
    s = scales.astype(np.uint8)
    sc = np.empty(s.shape[:-1] + (8,), np.uint8)
    mn = np.empty_like(sc)
    for j in range(8):
        if j < 4:
            sc[..., j] = s[..., j] & 63
            mn[..., j] = s[..., j + 4] & 63
        else:
            sc[..., j] = (s[..., j + 4] & 0x0F) | ((s[..., j - 4] >> 6) << 4)
            mn[..., j] = (s[..., j + 4] >> 4) | ((s[..., j] >> 6) << 4)
    return sc, mn


def dequant_q4k_f32(blocks: np.ndarray) -> np.ndarray:
    """[n,144] uint8 -> [n,256] float32, y = (d*sc)*q - (dmin*m) with each product and the difference
    rounded to fp32 separately (gguf/quants.py:504-521)."""
    b = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1, 144)
    n = b.shape[0]
    d = b[:, 0:2].copy().view(np.float16).astype(np.float32).reshape(n, 1)
    dmin = b[:, 2:4].copy().view(np.float16).astype(np.float32).reshape(n, 1)
    sc, mn = q4k_scale_min(b[:, 4:16])
    with np.errstate(all="ignore"):
        dsc = (d * sc.astype(np.float32)).astype(np.float32)   # [n,8]
        dmn = (dmin * mn.astype(np.float32)).astype(np.float32)
        qs = b[:, 16:144].reshape(n, 4, 32)
        q = np.empty((n, 8, 32), np.float32)
        q[:, 0::2, :] = (qs & 0x0F).astype(np.float32)
        q[:, 1::2, :] = (qs >> 4).astype(np.float32)
        prod = (dsc[:, :, None] * q).astype(np.float32)
        y = (prod - dmn[:, :, None]).astype(np.float32)
    return y.reshape(n, 256)


def dequant_q4k_bf16(blocks: np.ndarray) -> np.ndarray:
    return f32_to_bf16(dequant_q4k_f32(blocks)).reshape(-1, 256)


def dequant_q8_0_f32(blocks: np.ndarray) -> np.ndarray:
    """[n,34] uint8 (d f16 | 32 x int8) -> [n,32] float32, y = q * d in fp32 (gguf/quants.py:395-401)."""
    b = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1, 34)
    d = b[:, 0:2].copy().view(np.float16).astype(np.float32)
    q = b[:, 2:].view(np.int8).astype(np.float32)
    with np.errstate(all="ignore"):
        return (q * d).astype(np.float32)


def dequant_q8_0_bf16(blocks: np.ndarray) -> np.ndarray:
    return f32_to_bf16(dequant_q8_0_f32(blocks)).reshape(-1, 32)


def dequant_q6k_f32(blocks: np.ndarray) -> np.ndarray:
    """[n,210] uint8 (ql[128] | qh[64] | scales[16] int8 | d f16) -> [n,256] float32.  Explicit-index restatement of
    gguf/quants.py:552-572: element e = 32g+i takes its low nibble from ql[64*(g//4) + 32*(g%2) + i] >> 4*((g%4)//2),
    its high two bits from qh[32*(g//4) + i] >> 2*(g%4); q = (lo | hi<<4) - 32; y = (d*scales[e//16]) * q."""
    b = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1, 210)
    n = b.shape[0]
    ql, qh = b[:, 0:128], b[:, 128:192]
    sc = b[:, 192:208].view(np.int8).astype(np.float32)
    d = b[:, 208:210].copy().view(np.float16).astype(np.float32)
    q = np.empty((n, 8, 32), np.int16)
    for g in range(8):
        lo = (ql[:, 64 * (g // 4) + 32 * (g % 2): 64 * (g // 4) + 32 * (g % 2) + 32] >> (4 * ((g % 4) // 2))) & 0x0F
        hi = (qh[:, 32 * (g // 4): 32 * (g // 4) + 32] >> (2 * (g % 4))) & 0x03
        q[:, g, :] = (lo | (hi << 4)).astype(np.int16) - 32
    with np.errstate(all="ignore"):
        dsc = (d * sc).astype(np.float32)  # [n,16]
        y = (dsc[:, :, None] * q.reshape(n, 16, 16).astype(np.float32)).astype(np.float32)
    return y.reshape(n, 256)


def dequant_q6k_bf16(blocks: np.ndarray) -> np.ndarray:
    return f32_to_bf16(dequant_q6k_f32(blocks)).reshape(-1, 256)



def _f16(b: np.ndarray) -> np.ndarray:
    """[n,2] uint8 -> [n,1] float32 (exact widening of the little-endian fp16)."""
    return np.ascontiguousarray(b).view(np.float16).astype(np.float32).reshape(-1, 1)


def _nibbles_lo_then_hi(qs: np.ndarray) -> np.ndarray:
    """[n,16] packed bytes -> [n,32]: elements 0..15 are the low nibbles of qs[0..15], 16..31 the high nibbles
    (the 32-weight legacy blocks Q4_0/Q4_1/Q5_0/Q5_1; gguf/quants.py:227-228)."""
    return np.concatenate([qs & 0x0F, qs >> 4], axis=1)


def dequant_q4_0_f32(blocks: np.ndarray) -> np.ndarray:
    """[n,18] (d f16 | qs[16]) -> [n,32]: y = d * (q - 8)  (gguf/quants.py:220-231)."""
    b = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1, 18)
    q = _nibbles_lo_then_hi(b[:, 2:18]).astype(np.int8) - np.int8(8)
    with np.errstate(all="ignore"):
        return (_f16(b[:, 0:2]) * q.astype(np.float32)).astype(np.float32)


def dequant_q4_1_f32(blocks: np.ndarray) -> np.ndarray:
    """[n,20] (d f16 | m f16 | qs[16]) -> [n,32]: y = (d*q) + m, product and sum rounded separately (gguf/quants.py:254-267)."""
    b = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1, 20)
    q = _nibbles_lo_then_hi(b[:, 4:20]).astype(np.float32)
    with np.errstate(all="ignore"):
        return ((_f16(b[:, 0:2]) * q).astype(np.float32) + _f16(b[:, 2:4])).astype(np.float32)


def _q5_legacy_q(qh4: np.ndarray, qs: np.ndarray) -> np.ndarray:
    """5-bit values of a 32-weight block: low 4 bits as in Q4_0, bit 4 of element e is bit e of the LE u32 qh."""
    qh = np.ascontiguousarray(qh4).view("<u4").reshape(-1, 1)
    hi = ((qh >> np.arange(32, dtype=np.uint32).reshape(1, 32)) & np.uint32(1)).astype(np.uint8)
    return _nibbles_lo_then_hi(qs) | (hi << 4)


def dequant_q5_0_f32(blocks: np.ndarray) -> np.ndarray:
    """[n,22] (d f16 | qh u32 | qs[16]) -> [n,32]: y = d * (q5 - 16)  (gguf/quants.py:291-308)."""
    b = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1, 22)
    q = _q5_legacy_q(b[:, 2:6], b[:, 6:22]).astype(np.int8) - np.int8(16)
    with np.errstate(all="ignore"):
        return (_f16(b[:, 0:2]) * q.astype(np.float32)).astype(np.float32)


def dequant_q5_1_f32(blocks: np.ndarray) -> np.ndarray:
    """[n,24] (d f16 | m f16 | qh u32 | qs[16]) -> [n,32]: y = (d*q5) + m  (gguf/quants.py:333-352)."""
    b = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1, 24)
    q = _q5_legacy_q(b[:, 4:8], b[:, 8:24]).astype(np.float32)
    with np.errstate(all="ignore"):
        return ((_f16(b[:, 0:2]) * q).astype(np.float32) + _f16(b[:, 2:4])).astype(np.float32)


def dequant_q2k_f32(blocks: np.ndarray) -> np.ndarray:
    """[n,84] (scales[16] | qs[64] | d f16 | dmin f16) -> [n,256].  Element e = 128h + 32s + i (h<2, s<4, i<32) takes
    q = (qs[32h+i] >> 2s) & 3; its 16-weight sub-block j = e//16 has dl = d*(scales[j]&15), ml = dmin*(scales[j]>>4);
    y = dl*q - ml, every operation rounded to fp32 (gguf/quants.py:404-428)."""
    b = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1, 84)
    n = b.shape[0]
    sc, qs = b[:, 0:16], b[:, 16:80]
    d, dmin = _f16(b[:, 80:82]), _f16(b[:, 82:84])
    q = np.empty((n, 2, 4, 32), np.uint8)
    for h in range(2):
        for s in range(4):
            q[:, h, s, :] = (qs[:, 32 * h: 32 * h + 32] >> (2 * s)) & 3
    with np.errstate(all="ignore"):
        dl = (d * (sc & 0x0F).astype(np.float32)).astype(np.float32)
        ml = (dmin * (sc >> 4).astype(np.float32)).astype(np.float32)
        y = ((dl[:, :, None] * q.reshape(n, 16, 16).astype(np.float32)).astype(np.float32) - ml[:, :, None]).astype(np.float32)
    return y.reshape(n, 256)


def q3k_scales(s12: np.ndarray) -> np.ndarray:
    """12 packed bytes -> 16 signed 6-bit scales (gguf/quants.py:441-462): low 4 bits of scale k are s[k]&15 (k<8) or
    s[k-8]>>4 (k>=8); its high 2 bits are (s[8 + k%4] >> 2*(k//4)) & 3; value = (lo | hi<<4) - 32."""
    s = s12.astype(np.uint8)
    out = np.empty(s.shape[:-1] + (16,), np.int8)
    for k in range(16):
        lo = (s[..., k] & 0x0F) if k < 8 else (s[..., k - 8] >> 4)
        hi = (s[..., 8 + (k % 4)] >> (2 * (k // 4))) & 3
        out[..., k] = (lo | (hi << 4)).astype(np.int8) - np.int8(32)
    return out


def dequant_q3k_f32(blocks: np.ndarray) -> np.ndarray:
    """[n,110] (hmask[32] | qs[64] | scales[12] | d f16) -> [n,256].  Element e = 128h + 32s + i has low bits
    (qs[32h+i] >> 2s) & 3; written as e = 32b + i (b<8) its high bit is (hmask[i] >> b) & 1 and q = low - (4 if that
    bit is CLEAR else 0); y = (d*scale[e//16]) * q (gguf/quants.py:431-472)."""
    b = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1, 110)
    n = b.shape[0]
    hm, qs = b[:, 0:32], b[:, 32:96]
    sc = q3k_scales(b[:, 96:108]).astype(np.float32)
    d = _f16(b[:, 108:110])
    q = np.empty((n, 8, 32), np.int8)
    for g in range(8):  # g = 4h + s = e // 32
        lo = (qs[:, 32 * (g // 4): 32 * (g // 4) + 32] >> (2 * (g % 4))) & 3
        clear = ((hm >> g) & 1) ^ 1
        q[:, g, :] = lo.astype(np.int8) - (clear << 2).astype(np.int8)
    with np.errstate(all="ignore"):
        dl = (d * sc).astype(np.float32)
        y = (dl[:, :, None] * q.reshape(n, 16, 16).astype(np.float32)).astype(np.float32)
    return y.reshape(n, 256)


def dequant_q5k_f32(blocks: np.ndarray) -> np.ndarray:
    """[n,176] (d f16 | dmin f16 | scales[12] | qh[32] | qs[128]) -> [n,256].  Sub-block j (32 weights), element i:
    q = ((qs[32*(j//2)+i] >> 4*(j%2)) & 15) | (((qh[i] >> j) & 1) << 4); (scale, min) packed as in Q4_K;
    y = (d*sc_j)*q - (dmin*m_j) (gguf/quants.py:525-548)."""
    b = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1, 176)
    n = b.shape[0]
    d, dmin = _f16(b[:, 0:2]), _f16(b[:, 2:4])
    sc, mn = q4k_scale_min(b[:, 4:16])
    qh, qs = b[:, 16:48], b[:, 48:176]
    q = np.empty((n, 8, 32), np.uint8)
    for j in range(8):
        lo = (qs[:, 32 * (j // 2): 32 * (j // 2) + 32] >> (4 * (j % 2))) & 0x0F
        q[:, j, :] = lo | (((qh >> j) & 1) << 4)
    with np.errstate(all="ignore"):
        dsc = (d * sc.astype(np.float32)).astype(np.float32)
        dmn = (dmin * mn.astype(np.float32)).astype(np.float32)
        y = ((dsc[:, :, None] * q.astype(np.float32)).astype(np.float32) - dmn[:, :, None]).astype(np.float32)
    return y.reshape(n, 256)


IQ4NL_VALUES = np.array([-127, -104, -83, -65, -49, -35, -22, -10, 1, 13, 25, 38, 53, 69, 89, 113], np.int8)  # gguf/quants.py:1331
MXFP4_VALUES = np.array([0, 1, 2, 3, 4, 6, 8, 12, 0, -1, -2, -3, -4, -6, -8, -12], np.int8)  # e2m1 values, doubled (gguf/quants.py:659)


def dequant_iq4nl_f32(blocks: np.ndarray) -> np.ndarray:
    """[n,18] (d f16 | qs[16]) -> [n,32]: Q4_0's layout with a non-linear codebook, y = d * IQ4NL_VALUES[q4] (gguf/quants.py:1330-1348)."""
    b = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1, 18)
    kv = IQ4NL_VALUES[_nibbles_lo_then_hi(b[:, 2:18])].astype(np.float32)
    with np.errstate(all="ignore"):
        return (_f16(b[:, 0:2]) * kv).astype(np.float32)


def dequant_iq4xs_f32(blocks: np.ndarray) -> np.ndarray:
    """[n,136] (d f16 | scales_h u16 | scales_l[4] | qs[128]) -> [n,256].  Sub-block j (32 weights): 6-bit scale
    ls = ((scales_l[j//2] >> 4*(j%2)) & 15) | (((scales_h >> 2j) & 3) << 4), dl = d * (ls - 32); its elements i < 16 are the low
    nibbles of qs[16j + i], i >= 16 the high nibbles of qs[16j + i - 16]; y = dl * IQ4NL_VALUES[q4] (gguf/quants.py:1351-1380)."""
    b = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1, 136)
    n = b.shape[0]
    d = _f16(b[:, 0:2])
    sh = np.ascontiguousarray(b[:, 2:4]).view("<u2").reshape(n).astype(np.uint32)
    sl, qs = b[:, 4:8], b[:, 8:136]
    ls = np.empty((n, 8), np.uint8)
    q = np.empty((n, 8, 32), np.uint8)
    for j in range(8):
        ls[:, j] = ((sl[:, j // 2] >> (4 * (j % 2))) & 0x0F) | ((((sh >> (2 * j)) & 3) << 4).astype(np.uint8))
        q[:, j, :] = _nibbles_lo_then_hi(qs[:, 16 * j: 16 * j + 16])
    with np.errstate(all="ignore"):
        dl = (d * (ls.astype(np.int8) - np.int8(32)).astype(np.float32)).astype(np.float32)
        y = (dl[:, :, None] * IQ4NL_VALUES[q].astype(np.float32)).astype(np.float32)
    return y.reshape(n, 256)


def e8m0_to_f32_half(e: np.ndarray) -> np.ndarray:
    """E8M0 scale byte -> HALF the power of two it encodes, as fp32 (the codebook holds doubled e2m1 values):
    2^(e-128) — a subnormal for e < 2 (bits 0x00200000 << e), else exponent field e - 1 (ggml-impl.h ggml_e8m0_to_fp32_half)."""
    e = e.astype(np.uint32)
    return np.where(e < 2, np.uint32(0x00200000) << e, (e - np.uint32(1)) << np.uint32(23)).astype(np.uint32).view(np.float32)


def dequant_mxfp4_f32(blocks: np.ndarray) -> np.ndarray:
    """[n,17] (e u8 (E8M0) | qs[16]) -> [n,32]: y = e8m0_half(e) * MXFP4_VALUES[q4]; nibble order as in Q4_0 (gguf/quants.py:656-708)."""
    b = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1, 17)
    kv = MXFP4_VALUES[_nibbles_lo_then_hi(b[:, 1:17])].astype(np.float32)
    with np.errstate(all="ignore"):
        return (e8m0_to_f32_half(b[:, 0:1]) * kv).astype(np.float32)


# ---- lattice i-quants, ternary types, NVFP4 -------------------------------------------------------------------
# Codebooks: oracle/iq_grids.py (generated from the published ggml constants by tools/gen_iq_grids.py; pinned to gguf-py's in the tests).
def _grid(name: str) -> np.ndarray:
    from .iq_grids import GRIDS
    g = _GRID_CACHE.get(name)
    if g is None:
        g = _GRID_CACHE[name] = np.frombuffer(b"".join(GRIDS[name]), np.uint8).reshape(len(GRIDS[name]), -1).copy()
    return g


_GRID_CACHE: Dict[str, np.ndarray] = {}
# sign patterns of the IQ2/IQ3 "xxs/xs" types: 7 stored bits, the 8th is their parity (ggml ksigns_iq2xs)
KSIGNS = np.array([i | ((bin(i).count("1") & 1) << 7) for i in range(128)], np.uint8)


def _sign_from_bits(bits8: np.ndarray) -> np.ndarray:
    """[..] uint8 -> [.., 8] float32 of +1 / -1: bit k set means weight k is negated."""
    b = (bits8[..., None] >> np.arange(8, dtype=np.uint8)) & 1
    return np.where(b == 0, np.float32(1), np.float32(-1))


def dequant_iq2xxs_f32(blocks: np.ndarray) -> np.ndarray:
    """[n,66] (d f16 | 8 x (u32 idx4 | u32 signs+scale)) -> [n,256].  Group g (32 weights): four grid indices = the bytes of the first
    word; the second word holds four 7-bit sign-pattern indices and, in its top 4 bits, the scale s: db = d*(0.5+s)*0.25;
    y = (db * grid[idx][k]) * sign (gguf/quants.py IQ2_XXS)."""
    b = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1, 66)
    n = b.shape[0]
    d = _f16(b[:, 0:2])
    q = np.ascontiguousarray(b[:, 2:66]).view("<u4").reshape(n, 8, 2)
    idx = np.ascontiguousarray(q[:, :, 0]).view(np.uint8).reshape(n, 8, 4)
    sidx = (q[:, :, 1:2] >> np.array([0, 7, 14, 21], np.uint32)) & np.uint32(0x7F)
    with np.errstate(all="ignore"):
        db = ((d * (np.float32(0.5) + (q[:, :, 1] >> 28).astype(np.float32))).astype(np.float32) * np.float32(0.25)).astype(np.float32)
        y = ((db[:, :, None, None] * _grid("iq2xxs")[idx].astype(np.float32)).astype(np.float32) * _sign_from_bits(KSIGNS[sidx])).astype(np.float32)
    return y.reshape(n, 256)


def dequant_iq2xs_f32(blocks: np.ndarray) -> np.ndarray:
    """[n,74] (d f16 | 32 x u16 (9-bit grid index | 7-bit sign index << 9) | scales[8]) -> [n,256].  Entry j covers weights 8j..8j+7;
    scale of weights 16k..16k+15 is nibble k%2 of scales[k//2]: db = d*(0.5+s)*0.25 (gguf/quants.py IQ2_XS)."""
    b = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1, 74)
    n = b.shape[0]
    d = _f16(b[:, 0:2])
    q = np.ascontiguousarray(b[:, 2:66]).view("<u2").reshape(n, 32)
    sc = np.stack([b[:, 66:74] & 0x0F, b[:, 66:74] >> 4], axis=-1).reshape(n, 16)
    with np.errstate(all="ignore"):
        db = ((d * (np.float32(0.5) + sc.astype(np.float32))).astype(np.float32) * np.float32(0.25)).astype(np.float32)
        g = _grid("iq2xs")[q & 511].astype(np.float32).reshape(n, 16, 2, 8)
        sg = _sign_from_bits(KSIGNS[q >> 9]).reshape(n, 16, 2, 8)
        y = ((db[:, :, None, None] * g).astype(np.float32) * sg).astype(np.float32)
    return y.reshape(n, 256)


def dequant_iq2s_f32(blocks: np.ndarray) -> np.ndarray:
    """[n,82] (d f16 | qs[32] | signs[32] | qh[8] | scales[8]) -> [n,256].  Entry j: index qs[j] | ((qh[j//4] >> 2(j%4)) & 3) << 8,
    sign bits signs[j]; scales as IQ2_XS (gguf/quants.py IQ2_S)."""
    b = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1, 82)
    n = b.shape[0]
    d = _f16(b[:, 0:2])
    qs, signs, qh, scb = b[:, 2:34], b[:, 34:66], b[:, 66:74], b[:, 74:82]
    j = np.arange(32)
    idx = qs.astype(np.uint16) | (((qh[:, j // 4] >> (2 * (j % 4)).astype(np.uint8)) & 3).astype(np.uint16) << 8)
    sc = np.stack([scb & 0x0F, scb >> 4], axis=-1).reshape(n, 16)
    with np.errstate(all="ignore"):
        db = ((d * (np.float32(0.5) + sc.astype(np.float32))).astype(np.float32) * np.float32(0.25)).astype(np.float32)
        g = _grid("iq2s")[idx].astype(np.float32).reshape(n, 16, 2, 8)
        y = ((db[:, :, None, None] * g).astype(np.float32) * _sign_from_bits(signs).reshape(n, 16, 2, 8)).astype(np.float32)
    return y.reshape(n, 256)


def dequant_iq3xxs_f32(blocks: np.ndarray) -> np.ndarray:
    """[n,98] (d f16 | qs[64] | 8 x u32 signs+scale) -> [n,256].  Grid entries hold 4 values: weights 4j..4j+3 = grid[qs[j]]; group g
    (32 weights) has four 7-bit sign indices (8 weights each) and a 4-bit scale in its u32: db = d*(0.5+s)*0.5 (gguf/quants.py IQ3_XXS)."""
    b = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1, 98)
    n = b.shape[0]
    d = _f16(b[:, 0:2])
    qs = b[:, 2:66]
    w = np.ascontiguousarray(b[:, 66:98]).view("<u4").reshape(n, 8)
    sidx = (w[:, :, None] >> np.array([0, 7, 14, 21], np.uint32)) & np.uint32(0x7F)
    with np.errstate(all="ignore"):
        db = ((d * (np.float32(0.5) + (w >> 28).astype(np.float32))).astype(np.float32) * np.float32(0.5)).astype(np.float32)
        g = _grid("iq3xxs")[qs].astype(np.float32).reshape(n, 8, 4, 8)
        y = ((db[:, :, None, None] * g).astype(np.float32) * _sign_from_bits(KSIGNS[sidx])).astype(np.float32)
    return y.reshape(n, 256)


def dequant_iq3s_f32(blocks: np.ndarray) -> np.ndarray:
    """[n,110] (d f16 | qs[64] | qh[8] | signs[32] | scales[4]) -> [n,256].  Weights 4j..4j+3 = grid[qs[j] | ((qh[j//8] >> j%8) & 1) << 8];
    sign bits of weights 8m..8m+7 = signs[m]; scale of group g (32 weights) = nibble g%2 of scales[g//2]: db = d*(1+2s)
    (gguf/quants.py IQ3_S)."""
    b = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1, 110)
    n = b.shape[0]
    d = _f16(b[:, 0:2])
    qs, qh, signs, scb = b[:, 2:66], b[:, 66:74], b[:, 74:106], b[:, 106:110]
    j = np.arange(64)
    idx = qs.astype(np.uint16) | (((qh[:, j // 8] >> (j % 8).astype(np.uint8)) & 1).astype(np.uint16) << 8)
    sc = np.stack([scb & 0x0F, scb >> 4], axis=-1).reshape(n, 8)
    with np.errstate(all="ignore"):
        db = (d * (1 + 2 * sc).astype(np.float32)).astype(np.float32)
        g = _grid("iq3s")[idx].astype(np.float32).reshape(n, 8, 4, 8)
        y = ((db[:, :, None, None] * g).astype(np.float32) * _sign_from_bits(signs).reshape(n, 8, 4, 8)).astype(np.float32)
    return y.reshape(n, 256)


IQ1_DELTA = np.float32(0.125)


def dequant_iq1s_f32(blocks: np.ndarray) -> np.ndarray:
    """[n,50] (d f16 | qs[32] | 8 x u16 qh) -> [n,256].  Group g (32 weights), entry k<4 (8 weights): 11-bit index
    qs[4g+k] | ((qh[g] >> 3k) & 7) << 8 into a grid of values in {-1,0,1}; dl = d*(2*((qh[g]>>12)&7)+1); delta = -0.125 when bit 15 of
    qh[g] is set else +0.125; y = dl * (grid + delta) (gguf/quants.py IQ1_S)."""
    b = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1, 50)
    n = b.shape[0]
    d = _f16(b[:, 0:2])
    qs = b[:, 2:34].reshape(n, 8, 4)
    qh = np.ascontiguousarray(b[:, 34:50]).view("<u2").reshape(n, 8)
    idx = qs.astype(np.uint16) | (((qh[:, :, None] >> np.array([0, 3, 6, 9], np.uint16)) & 7) << 8)
    with np.errstate(all="ignore"):
        dl = (d * (2 * ((qh >> 12) & 7) + 1).astype(np.float32)).astype(np.float32)
        delta = np.where((qh & 0x8000) == 0, IQ1_DELTA, -IQ1_DELTA).astype(np.float32)
        g = _grid("iq1s")[idx].astype(np.float32) - np.float32(1)  # stored + 1
        y = (dl[:, :, None, None] * (g + delta[:, :, None, None]).astype(np.float32)).astype(np.float32)
    return y.reshape(n, 256)


def dequant_iq1m_f32(blocks: np.ndarray) -> np.ndarray:
    """[n,56] (qs[32] | qh[16] | 4 x u16 scales) -> [n,256].  The fp16 super-scale d is spread over the top nibbles of the four scale
    words (word i holds bits 4i..4i+3); 3-bit scale of weights 16k..16k+15 = (scales16[k//4] >> 3(k%4)) & 7, dl = d*(2s+1).  Entry j
    (8 weights): nibble j%2 of qh[j//2]: low 3 bits extend the index qs[j] to 11 bits (grid shared with IQ1_S), bit 3 selects
    delta = -0.125 (gguf/quants.py IQ1_M)."""
    b = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1, 56)
    n = b.shape[0]
    qs, qh = b[:, 0:32], b[:, 32:48]
    sc16 = np.ascontiguousarray(b[:, 48:56]).view("<u2").reshape(n, 4)
    dbits = ((sc16[:, 0] >> 12) | ((sc16[:, 1] >> 12) << 4) | ((sc16[:, 2] >> 12) << 8) | ((sc16[:, 3] >> 12) << 12)).astype(np.uint16)
    d = dbits.view(np.float16).astype(np.float32).reshape(n, 1)
    k = np.arange(16)
    s3 = (sc16[:, k // 4] >> (3 * (k % 4)).astype(np.uint16)) & 7
    j = np.arange(32)
    nib = (qh[:, j // 2] >> (4 * (j % 2)).astype(np.uint8)) & 0x0F
    idx = qs.astype(np.uint16) | ((nib & 7).astype(np.uint16) << 8)
    with np.errstate(all="ignore"):
        dl = (d * (2 * s3 + 1).astype(np.float32)).astype(np.float32)            # [n,16]
        delta = np.where((nib & 8) == 0, IQ1_DELTA, -IQ1_DELTA).astype(np.float32)  # [n,32]
        g = _grid("iq1s")[idx].astype(np.float32) - np.float32(1)
        y = (np.repeat(dl, 2, axis=1)[:, :, None] * (g + delta[:, :, None]).astype(np.float32)).astype(np.float32)
    return y.reshape(n, 256)


def dequant_tq2_0_f32(blocks: np.ndarray) -> np.ndarray:
    """[n,66] (qs[64] | d f16) -> [n,256]: weight 128h + 32s + i = ((qs[32h+i] >> 2s) & 3) - 1, times d (gguf/quants.py TQ2_0)."""
    b = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1, 66)
    n = b.shape[0]
    q = np.empty((n, 2, 4, 32), np.int8)
    for h in range(2):
        for s in range(4):
            q[:, h, s, :] = ((b[:, 32 * h: 32 * h + 32] >> (2 * s)) & 3).astype(np.int8) - np.int8(1)
    with np.errstate(all="ignore"):
        return (_f16(b[:, 64:66]) * q.reshape(n, 256).astype(np.float32)).astype(np.float32)


def dequant_tq1_0_f32(blocks: np.ndarray) -> np.ndarray:
    """[n,54] (qs[48] | qh[4] | d f16) -> [n,256].  Base-3 digits, five per byte (four per qh byte): weight e takes byte B and power p,
    v = (B * 3^p) mod 256, trit = (v * 3) >> 8, y = d * (trit - 1).  e < 160: B = qs[e%32], p = e//32; 160 <= e < 240: B = qs[32 + (e-160)%16],
    p = (e-160)//16; e >= 240: B = qh[(e-240)%4], p = (e-240)//4 (gguf/quants.py TQ1_0)."""
    b = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1, 54)
    n = b.shape[0]
    e = np.arange(256)
    byte = np.where(e < 160, e % 32, np.where(e < 240, 32 + (e - 160) % 16, 48 + (e - 240) % 4))
    powr = np.where(e < 160, e // 32, np.where(e < 240, (e - 160) // 16, (e - 240) // 4))
    v = (b[:, byte].astype(np.uint16) * (3 ** powr).astype(np.uint16)) & 0xFF
    trit = ((v * 3) >> 8).astype(np.int8) - np.int8(1)
    with np.errstate(all="ignore"):
        return (_f16(b[:, 52:54]) * trit.astype(np.float32)).astype(np.float32)


NVFP4_VALUES = MXFP4_VALUES  # the same doubled e2m1 code points


def ue4m3_to_f32_half(x: np.ndarray) -> np.ndarray:
    """Unsigned E4M3 scale byte (bias 7; bit 7 ignored) -> HALF its value as fp32; 0x00 and 0x7F decode to 0 (gguf/quants.py NVFP4.ue4m3_to_fp32)."""
    x = x.astype(np.uint8)
    e = ((x >> 3) & 0xF).astype(np.int32)
    m = (x & 7).astype(np.float32)
    raw = np.where(e == 0, m * np.float32(2.0 ** -9), (np.float32(1) + m / np.float32(8)) * np.exp2(e.astype(np.float32) - np.float32(7)))
    return np.where((x == 0) | (x == 0x7F), np.float32(0), raw.astype(np.float32) * np.float32(0.5)).astype(np.float32)


def dequant_nvfp4_f32(blocks: np.ndarray) -> np.ndarray:
    """[n,36] (4 x UE4M3 scale | qs[32]) -> [n,64]: sub-block s (16 weights) uses scale byte s and qs[8s..8s+7]: its first 8 weights are the
    low nibbles, the next 8 the high nibbles; y = scale_half * NVFP4_VALUES[q4] (gguf/quants.py NVFP4)."""
    b = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1, 36)
    n = b.shape[0]
    qs = b[:, 4:36].reshape(n, 4, 8)
    q = np.concatenate([qs & 0x0F, qs >> 4], axis=-1)
    with np.errstate(all="ignore"):
        return (ue4m3_to_f32_half(b[:, 0:4])[:, :, None] * NVFP4_VALUES[q].astype(np.float32)).astype(np.float32).reshape(n, 64)


# file dtype -> (weights per block, bytes per block, fp32 dequantiser).  Everything the product dequantises to bf16.
BLOCK_QUANTS = {
    "Q4_0": (32, 18, dequant_q4_0_f32), "Q4_1": (32, 20, dequant_q4_1_f32), "Q5_0": (32, 22, dequant_q5_0_f32),
    "Q5_1": (32, 24, dequant_q5_1_f32), "Q8_0": (32, 34, dequant_q8_0_f32), "Q2_K": (256, 84, dequant_q2k_f32),
    "Q3_K": (256, 110, dequant_q3k_f32), "Q4_K": (256, 144, dequant_q4k_f32), "Q5_K": (256, 176, dequant_q5k_f32),
    "Q6_K": (256, 210, dequant_q6k_f32), "IQ4_NL": (32, 18, dequant_iq4nl_f32), "IQ4_XS": (256, 136, dequant_iq4xs_f32),
    "MXFP4": (32, 17, dequant_mxfp4_f32), "IQ2_XXS": (256, 66, dequant_iq2xxs_f32), "IQ2_XS": (256, 74, dequant_iq2xs_f32),
    "IQ2_S": (256, 82, dequant_iq2s_f32), "IQ3_XXS": (256, 98, dequant_iq3xxs_f32), "IQ3_S": (256, 110, dequant_iq3s_f32),
    "IQ1_S": (256, 50, dequant_iq1s_f32), "IQ1_M": (256, 56, dequant_iq1m_f32), "TQ1_0": (256, 54, dequant_tq1_0_f32),
    "TQ2_0": (256, 66, dequant_tq2_0_f32), "NVFP4": (64, 36, dequant_nvfp4_f32),
}


def dequant_bf16(dtype: str, blocks: np.ndarray) -> np.ndarray:
    """Block-quantised bytes of `dtype` -> bf16 bit patterns [n_blocks, weights per block]."""
    nel, nb, fn = BLOCK_QUANTS[dtype]
    return f32_to_bf16(fn(np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1, nb))).reshape(-1, nel)


_MASK = (1 << 64) - 1


def checksum(data: bytes | np.ndarray) -> int:
    """kk_checksum: sum_i mix(w_i + i*0x9E3779B97F4A7C15) mod 2^64 over little-endian 8-byte words
    (tail zero-padded); mix = splitmix64 finaliser."""
    a = np.frombuffer(bytes(data) if not isinstance(data, np.ndarray) else data.tobytes(), dtype=np.uint8)
    pad = (-len(a)) % 8
    if pad:
        a = np.concatenate([a, np.zeros(pad, np.uint8)])
    w = a.view("<u8").astype(np.uint64)
    i = np.arange(len(w), dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = w + i * np.uint64(0x9E3779B97F4A7C15)
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
        return int(np.sum(x, dtype=np.uint64))


# ---------------------------------------------------------------------------------------------
# pool layout + expected pool contents ("Load")
# ---------------------------------------------------------------------------------------------
GPT2_CONV1D = ("attn.c_attn.weight", "attn.c_proj.weight", "mlp.c_fc.weight", "mlp.c_proj.weight")
SCATTER_DIM0 = ("q_proj.weight", "k_proj.weight", "v_proj.weight", "gate_proj.weight", "up_proj.weight",
                "embed_tokens.weight", "lm_head.weight", "attn_q.weight", "attn_k.weight", "attn_v.weight",
                "ffn_gate.weight", "ffn_up.weight", "token_embd.weight", "output.weight")
SCATTER_DIM1 = ("o_proj.weight", "down_proj.weight", "attn_output.weight", "ffn_down.weight")
MODE_SINGLE, MODE_BROADCAST, MODE_SCATTER = 0, 1, 2
LOAD_GPT2_CONV1D_T, LOAD_KEEP_F32, LOAD_F8_TO_BF16 = 0x1, 0x2, 0x10
_ELEM_BYTES = {k: v // 8 for k, v in ST_BITS.items() if v % 8 == 0}
_ELEM_BYTES.update({"F32": 4, "F16": 2, "BF16": 2, "I8": 1, "I16": 2, "I32": 4, "I64": 8, "F64": 8})


def pool_dtype(dt: str, flags: int = 0) -> str:
    if dt in ("F16", "BF16") or dt in BLOCK_QUANTS:
        return "BF16"
    if dt == "F32":
        return "F32" if flags & LOAD_KEEP_F32 else "BF16"
    if dt in ("F8_E4M3", "F8_E5M2") and flags & LOAD_F8_TO_BF16:
        return "BF16"
    return dt


def slice_dim(rec: dict, n_parts: int):
    """[PROPOSED] scatter rule: column-parallel weights sliced along dim 0,
    row-parallel along dim 1, everything else (1-D, >2-D, indivisible, dim-1 of block-quantised) whole."""
    if n_parts <= 1 or len(rec["shape"]) != 2:
        return None
    n = rec["name"]
    dim = None
    if any(n.endswith(s) for s in SCATTER_DIM1):
        dim = 1
    elif any(n.endswith(s) for s in SCATTER_DIM0) and "norm" not in n:
        dim = 0
    if dim is None or 0 in rec["shape"] or rec["shape"][dim] % n_parts:
        return None
    if rec["dtype"] in ("F4", "F6_E2M3", "F6_E3M2"):
        return None
    if rec["dtype"] in BLOCK_QUANTS and dim == 1:
        return None
    return dim


def plan_pool(recs: List[dict], mode: int = MODE_SINGLE, flags: int = 0, n_parts: int = 1, part: int = 0) -> Tuple[List[dict], int]:
    """Pool placement of every tensor for one device: index order, 256-byte aligned slots."""
    off = 0
    out = []
    for r in recs:
        pdt = pool_dtype(r["dtype"], flags)
        shape = list(r["shape"])
        p = dict(name=r["name"], dtype=pdt, slice_dim=None, slice_begin=0)
        tr = bool(flags & LOAD_GPT2_CONV1D_T) and mode != MODE_SCATTER and len(shape) == 2 and \
            any(r["name"].endswith(s) for s in GPT2_CONV1D) and r["dtype"] in ("F32", "F16", "BF16", "I16", "U16", "I32", "U32")
        if tr:
            shape = [shape[1], shape[0]]
        sd = slice_dim(r, n_parts) if mode == MODE_SCATTER else None
        if sd is not None:
            k = shape[sd] // n_parts
            p["slice_dim"], p["slice_begin"] = sd, k * part
            shape[sd] = k
        nel = 1
        for s in shape:
            nel *= s
        if r["dtype"] in BLOCK_QUANTS:
            nbytes = nel * 2
        elif pdt in _ELEM_BYTES:
            nbytes = nel * _ELEM_BYTES[pdt]
        else:  # sub-byte verbatim
            nbytes = r["nbytes"]
        p.update(shape=shape, transposed=tr, pool_offset=off, nbytes=nbytes)
        out.append(p)
        off = (off + nbytes + POOL_ALIGN - 1) // POOL_ALIGN * POOL_ALIGN
    return out, (off or POOL_ALIGN)


def convert_tensor(rec: dict, raw: bytes, flags: int = 0) -> np.ndarray:
    """File bytes of one tensor -> pool bytes (uint8 array), before any slicing/transposition."""
    dt = rec["dtype"]
    a = np.frombuffer(raw, dtype=np.uint8)
    if dt == "F32" and not (flags & LOAD_KEEP_F32):
        return f32_bits_to_bf16(a.view("<u4")).view(np.uint8)
    if dt == "F16":
        return f16_bits_to_bf16(a.view("<u2")).view(np.uint8)
    if dt == "F8_E4M3" and flags & LOAD_F8_TO_BF16:
        return f8e4m3_bits_to_bf16(a).view(np.uint8)
    if dt == "F8_E5M2" and flags & LOAD_F8_TO_BF16:
        return f8e5m2_bits_to_bf16(a).view(np.uint8)
    if dt in BLOCK_QUANTS:
        return dequant_bf16(dt, a).reshape(-1).view(np.uint8)
    if dt in ("Q8_K", "Q8_1"):
        raise OracleError(f"{dt} dequantisation not defined by this oracle")
    return a.copy()


def expected_pool(shards: List[str], recs: List[dict], mode: int = MODE_SINGLE, flags: int = 0,
                  n_parts: int = 1, part: int = 0) -> Tuple[np.ndarray, List[dict]]:
    """Whole expected pool of one device as a uint8 array (gaps zero).  Small checkpoints only."""
    plan, total = plan_pool(recs, mode, flags, n_parts, part)
    pool = np.zeros(total, np.uint8)
    fhs = [open(s, "rb") for s in shards]
    try:
        for r, p in zip(recs, plan):
            fh = fhs[r["shard"]]
            fh.seek(r["file_offset"])
            raw = fh.read(r["nbytes"])
            assert len(raw) == r["nbytes"]
            out = convert_tensor(r, raw, flags)
            es = out.size // max(1, int(np.prod(r["shape"]))) if int(np.prod(r["shape"])) else 1
            if p["transposed"]:
                R, C = r["shape"]
                out = np.ascontiguousarray(out.reshape(R, C, es).transpose(1, 0, 2)).reshape(-1)
            elif p["slice_dim"] is not None:
                R, C = r["shape"]
                v = out.reshape(R, C, es)
                k = p["shape"][p["slice_dim"]]
                b = p["slice_begin"]
                v = v[b:b + k] if p["slice_dim"] == 0 else v[:, b:b + k]
                out = np.ascontiguousarray(v).reshape(-1)
            assert out.size == p["nbytes"], (r["name"], out.size, p["nbytes"])
            pool[p["pool_offset"]: p["pool_offset"] + out.size] = out
    finally:
        for fh in fhs:
            fh.close()
    return pool, plan
