"""ctypes front-end of oracle/kk_oracle.c (TEST INFRASTRUCTURE ONLY — see oracle/oracle.py header)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libkk_oracle.so")
_lib = None

OP_COPY, OP_F32_BF16, OP_F16_BF16, OP_Q4K_BF16, OP_Q8_0_BF16, OP_Q6K_BF16 = 0, 1, 2, 3, 4, 5
OP_F8E4M3_BF16, OP_F8E5M2_BF16 = 6, 7
OP_DEQUANT = 0x100  # | ggml type id
# file dtype -> (ggml type id, weights per block, bytes per block)   (gguf/constants.py GGML_QUANT_SIZES)
GGML_BLOCK = {"Q4_0": (2, 32, 18), "Q4_1": (3, 32, 20), "Q5_0": (6, 32, 22), "Q5_1": (7, 32, 24), "Q8_0": (8, 32, 34),
              "Q2_K": (10, 256, 84), "Q3_K": (11, 256, 110), "Q4_K": (12, 256, 144), "Q5_K": (13, 256, 176), "Q6_K": (14, 256, 210),
              "IQ4_NL": (20, 32, 18), "IQ4_XS": (23, 256, 136), "MXFP4": (39, 32, 17), "IQ2_XXS": (16, 256, 66), "IQ2_XS": (17, 256, 74),
              "IQ3_XXS": (18, 256, 98), "IQ1_S": (19, 256, 50), "IQ3_S": (21, 256, 110), "IQ2_S": (22, 256, 82), "IQ1_M": (29, 256, 56),
              "TQ1_0": (34, 256, 54), "TQ2_0": (35, 256, 66), "NVFP4": (40, 64, 36)}


class OrcJob(C.Structure):
    _fields_ = [("shard", C.c_uint32), ("op", C.c_uint32), ("file_off", C.c_uint64), ("nbytes", C.c_uint64), ("dst_off", C.c_uint64)]


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "kk_oracle.c")
    newest = max(os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "iq_grids_c.h")))
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < newest:
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        u8p, u16p, u32p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint16), C.POINTER(C.c_uint32)
        L.orc_f32_to_bf16.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.orc_f16_to_bf16.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.orc_q4k_to_bf16.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.orc_q8_0_to_bf16.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.orc_q6k_to_bf16.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.orc_f8_to_bf16.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint64]
        L.orc_dequant_to_bf16.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64]
        L.orc_dequant_to_bf16.restype = C.c_int
        L.orc_checksum.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_checksum.restype = C.c_uint64
        for fn in ("orc_fill_bytes", "orc_fill_bf16_finite", "orc_fill_f32", "orc_fill_q4k"):
            getattr(L, fn).argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
            getattr(L, fn).restype = None
        L.orc_cpu_load.argtypes = [C.POINTER(C.c_char_p), C.c_uint32, C.POINTER(OrcJob), C.c_uint64, C.c_void_p, C.c_uint64, C.c_int]
        L.orc_cpu_load.restype = C.c_int
        L.orc_max_threads.restype = C.c_int
        _lib = L
        del u8p, u16p, u32p
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def f32_to_bf16(u32: np.ndarray) -> np.ndarray:
    src = np.ascontiguousarray(u32).view(np.uint32).reshape(-1)
    out = np.empty(src.size, np.uint16)
    lib().orc_f32_to_bf16(_ptr(src), _ptr(out), src.size)
    return out


def f16_to_bf16(u16: np.ndarray) -> np.ndarray:
    src = np.ascontiguousarray(u16).view(np.uint16).reshape(-1)
    out = np.empty(src.size, np.uint16)
    lib().orc_f16_to_bf16(_ptr(src), _ptr(out), src.size)
    return out


def q4k_to_bf16(blocks: np.ndarray) -> np.ndarray:
    src = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1)
    assert src.size % 144 == 0
    n = src.size // 144
    out = np.empty(n * 256, np.uint16)
    lib().orc_q4k_to_bf16(_ptr(src), _ptr(out), n)
    return out.reshape(n, 256)


def q8_0_to_bf16(blocks: np.ndarray) -> np.ndarray:
    src = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1)
    assert src.size % 34 == 0
    n = src.size // 34
    out = np.empty(n * 32, np.uint16)
    lib().orc_q8_0_to_bf16(_ptr(src), _ptr(out), n)
    return out.reshape(n, 32)


def q6k_to_bf16(blocks: np.ndarray) -> np.ndarray:
    src = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1)
    assert src.size % 210 == 0
    n = src.size // 210
    out = np.empty(n * 256, np.uint16)
    lib().orc_q6k_to_bf16(_ptr(src), _ptr(out), n)
    return out.reshape(n, 256)


def f8_to_bf16(dtype: str, u8: np.ndarray) -> np.ndarray:
    src = np.ascontiguousarray(u8, dtype=np.uint8).reshape(-1)
    out = np.empty(src.size, np.uint16)
    lib().orc_f8_to_bf16({"F8_E4M3": 0, "F8_E5M2": 1}[dtype], _ptr(src), _ptr(out), src.size)
    return out


def dequant_to_bf16(dtype: str, blocks: np.ndarray) -> np.ndarray:
    """Any block-quantised GGUF type the oracle defines -> bf16 bit patterns [n_blocks, weights per block]."""
    tid, nel, nb = GGML_BLOCK[dtype]
    src = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1)
    assert src.size % nb == 0
    n = src.size // nb
    out = np.empty(n * nel, np.uint16)
    if lib().orc_dequant_to_bf16(tid, _ptr(src), _ptr(out), n) != 0:
        raise ValueError(dtype)
    return out.reshape(n, nel)


def checksum(a: np.ndarray) -> int:
    src = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
    return int(lib().orc_checksum(_ptr(src), src.size))


def fill(kind: str, n: int, seed: int) -> np.ndarray:
    """kind: 'bytes' (n bytes), 'bf16' (n elems), 'f32' (n elems), 'q4k' (n blocks)."""
    L = lib()
    if kind == "bytes":
        a = np.empty(n, np.uint8); L.orc_fill_bytes(_ptr(a), n, seed)
    elif kind == "bf16":
        a = np.empty(n, np.uint16); L.orc_fill_bf16_finite(_ptr(a), n, seed)
    elif kind == "f32":
        a = np.empty(n, np.float32); L.orc_fill_f32(_ptr(a), n, seed)
    elif kind == "q4k":
        a = np.empty(n * 144, np.uint8); L.orc_fill_q4k(_ptr(a), n, seed)
    else:
        raise ValueError(kind)
    return a


def fill_into(kind: str, dst: np.ndarray, seed: int) -> None:
    L = lib()
    d = dst.reshape(-1)
    if kind == "bytes":
        L.orc_fill_bytes(_ptr(d), d.view(np.uint8).size, seed)
    elif kind == "bf16":
        L.orc_fill_bf16_finite(_ptr(d), d.view(np.uint16).size, seed)
    elif kind == "f32":
        L.orc_fill_f32(_ptr(d), d.view(np.float32).size, seed)
    elif kind == "q4k":
        L.orc_fill_q4k(_ptr(d), d.view(np.uint8).size // 144, seed)
    else:
        raise ValueError(kind)


_OPS = {"BF16": OP_COPY, "F32": OP_F32_BF16, "F16": OP_F16_BF16, "Q4_K": OP_Q4K_BF16, "Q8_0": OP_Q8_0_BF16, "Q6_K": OP_Q6K_BF16}
_OPS.update({dt: OP_DEQUANT | tid for dt, (tid, _, _) in GGML_BLOCK.items() if dt not in _OPS})
_UNITS = {OP_COPY: (1, 1), OP_F32_BF16: (4, 2), OP_F16_BF16: (2, 2)}  # op -> (source bytes, pool bytes) per unit
_UNITS.update({OP_F8E4M3_BF16: (1, 2), OP_F8E5M2_BF16: (1, 2)})
_UNITS.update({_OPS[dt]: (nb, 2 * nel) for dt, (_, nel, nb) in GGML_BLOCK.items()})


def make_jobs(recs: Sequence[dict], plan: Sequence[dict], job_bytes: int = 8 << 20, max_src_bytes: int | None = None):
    """Split every tensor into <= job_bytes pieces on unit boundaries (non-transposed, unsliced plans)."""
    jobs: List[OrcJob] = []
    total = 0
    for r, p in zip(recs, plan):
        if p["dtype"] == r["dtype"]:  # kept verbatim (BF16, integers, F32 under KEEP_F32, FP8 without F8_TO_BF16)
            op = OP_COPY
        elif r["dtype"] in ("F8_E4M3", "F8_E5M2"):
            op = OP_F8E4M3_BF16 if r["dtype"] == "F8_E4M3" else OP_F8E5M2_BF16
        else:
            op = _OPS.get(r["dtype"], OP_COPY)
        unit, out_unit = _UNITS[op]
        step = max(unit, job_bytes // unit * unit)
        off = 0
        while off < r["nbytes"]:
            n = min(step, r["nbytes"] - off)
            if max_src_bytes is not None and total + n > max_src_bytes:
                return jobs, total
            jobs.append(OrcJob(r["shard"], op, r["file_offset"] + off, n, p["pool_offset"] + off // unit * out_unit))
            total += n
            off += n
    return jobs, total


def cpu_load(shards: Sequence[str], jobs: Sequence[OrcJob], pool: np.ndarray, threads: int = 0, scratch: int = 16 << 20) -> None:
    arr = (OrcJob * len(jobs))(*jobs)
    paths = (C.c_char_p * len(shards))(*[s.encode() for s in shards])
    rc = lib().orc_cpu_load(paths, len(shards), arr, len(jobs), _ptr(pool), scratch, threads)
    if rc != 0:
        raise OSError("orc_cpu_load failed")


def max_threads() -> int:
    return int(lib().orc_max_threads())
