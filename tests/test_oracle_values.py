"""Pins the oracle's value arithmetic: numpy restatement == C twin == committed golden vectors ==
format owners' libraries (torch RNE, gguf-py Q4_K) run live."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def bits16(t):
    return t.contiguous().view(torch.int16).numpy().view(np.uint16)


def test_f32_golden_vectors(coracle):
    v = np.load(os.path.join(G, "cast_vectors.npz"))
    assert (oracle.f32_bits_to_bf16(v["f32_in"]) == v["f32_out"]).all()
    assert (coracle.f32_to_bf16(v["f32_in"]) == v["f32_out"]).all()


def test_f16_golden_vectors_every_finite_half(coracle):
    v = np.load(os.path.join(G, "cast_vectors.npz"))
    assert len(v["f16_in"]) > 63000
    assert (oracle.f16_bits_to_bf16(v["f16_in"]) == v["f16_out"]).all()
    assert (coracle.f16_to_bf16(v["f16_in"]) == v["f16_out"]).all()


def test_f32_random_vs_torch_live(coracle):
    rng = np.random.default_rng(3)
    u = rng.integers(0, 1 << 32, size=200_000, dtype=np.uint64).astype(np.uint32)
    u = u[(u & 0x7FFFFFFF) <= 0x7F800000]  # NaN handling is implementation-defined; pinned separately below
    ref = bits16(torch.from_numpy(u.view(np.float32).copy()).to(torch.bfloat16))
    assert (oracle.f32_bits_to_bf16(u) == ref).all()
    assert (coracle.f32_to_bf16(u) == ref).all()


def test_nan_rule_is_canonical_7fff(coracle):
    u = np.array([0x7FC00000, 0xFFC00000, 0x7F800001, 0xFFFFFFFF, 0x7FBFFFFF], np.uint32)
    assert (oracle.f32_bits_to_bf16(u) == 0x7FFF).all()
    assert (coracle.f32_to_bf16(u) == 0x7FFF).all()
    h = np.array([0x7E00, 0xFE00, 0x7C01], np.uint16)
    assert (oracle.f16_bits_to_bf16(h) == 0x7FFF).all()
    assert (coracle.f16_to_bf16(h) == 0x7FFF).all()


def test_q4k_golden_file_vs_gguf_py():
    import json
    exp = json.load(open(os.path.join(G, "q4k.gguf.expected.json")))
    outs = np.load(os.path.join(G, "q4k.gguf.bf16.npz"))
    raw = open(os.path.join(G, "q4k.gguf"), "rb").read()
    n = 0
    for t in exp["tensors"]:
        if t["dtype"] != "Q4_K":
            continue
        blocks = np.frombuffer(raw[t["file_offset"]:t["file_offset"] + t["nbytes"]], np.uint8).reshape(-1, 144)
        assert (oracle.dequant_q4k_bf16(blocks).reshape(-1) == outs[t["name"]]).all()
        n += 1
    assert n == 2


def test_q4k_random_blocks_vs_gguf_py_live_bit_exact(coracle):
    from gguf import GGMLQuantizationType, quants
    q = coracle.fill("q4k", 4096, 12345).reshape(-1, 144)
    ref = bits16(torch.from_numpy(quants.dequantize(q, GGMLQuantizationType.Q4_K)).to(torch.bfloat16))
    assert (oracle.dequant_q4k_bf16(q) == ref).all()
    assert (coracle.q4k_to_bf16(q) == ref).all()


def test_q4k_scale_min_unpack_vs_gguf_py():
    from gguf.quants import Q4_K
    rng = np.random.default_rng(5)
    s = rng.integers(0, 256, size=(1000, 12), dtype=np.uint8)
    sc, mn = oracle.q4k_scale_min(s)
    rsc, rmn = Q4_K.get_scale_min(s)
    assert (sc == rsc).all() and (mn == rmn).all()


def test_q4k_extreme_scales(coracle):
    # all-ones scales (63/63), zero d, max nibbles: exercises every bit of the 6-bit unpack
    b = np.zeros((3, 144), np.uint8)
    b[0, 4:16] = 0xFF; b[0, 16:] = 0xFF; b[0, 0:4] = np.array([0x3C00, 0x3C00], "<u2").view(np.uint8)  # d=dmin=1
    b[1, 4:16] = 0xAA; b[1, 16:] = 0x5A; b[1, 0:4] = np.array([0x0400, 0x0001], "<u2").view(np.uint8)  # tiny / subnormal
    b[2, 4:16] = 0x3F; b[2, 16:] = 0xF0
    from gguf import GGMLQuantizationType, quants
    ref = bits16(torch.from_numpy(quants.dequantize(b, GGMLQuantizationType.Q4_K)).to(torch.bfloat16))
    assert (oracle.dequant_q4k_bf16(b) == ref).all()
    assert (coracle.q4k_to_bf16(b) == ref).all()


def test_checksum_c_matches_numpy_and_is_position_sensitive(coracle):
    rng = np.random.default_rng(9)
    for n in (0, 1, 7, 8, 9, 4096, 100_003):
        a = rng.integers(0, 256, size=n, dtype=np.uint8)
        assert coracle.checksum(a) == oracle.checksum(a)
    a = rng.integers(0, 256, size=64, dtype=np.uint8)
    b = a.copy(); b[:8], b[8:16] = a[8:16].copy(), a[:8].copy()
    assert oracle.checksum(a) != oracle.checksum(b)


def test_fill_generators_are_finite_and_reproducible(coracle):
    a = coracle.fill("bf16", 100_000, 1)
    assert ((a & 0x7F80) != 0x7F80).all()
    assert (a == coracle.fill("bf16", 100_000, 1)).all() and (a != coracle.fill("bf16", 100_000, 2)).any()
    q = coracle.fill("q4k", 100, 1).reshape(-1, 144)
    d = q[:, 0:4].copy().view(np.float16).astype(np.float32)
    assert np.isfinite(d).all() and (d >= 2.0 ** -10).all() and (d < 2.0 ** -3).all()


def test_q6k_q8_0_random_blocks_vs_gguf_py_live_bit_exact(coracle):
    from gguf import GGMLQuantizationType, quants
    from tools import synth
    q8 = synth.gen_bytes("Q8_0", 34 * 6000, 3, 1).reshape(-1, 34)
    ref = bits16(torch.from_numpy(quants.dequantize(q8, GGMLQuantizationType.Q8_0)).to(torch.bfloat16))
    assert (oracle.dequant_q8_0_bf16(q8) == ref).all() and (coracle.q8_0_to_bf16(q8) == ref).all()
    q6 = synth.gen_bytes("Q6_K", 210 * 3000, 3, 2).reshape(-1, 210)
    ref = bits16(torch.from_numpy(quants.dequantize(q6, GGMLQuantizationType.Q6_K)).to(torch.bfloat16))
    assert (oracle.dequant_q6k_bf16(q6) == ref).all() and (coracle.q6k_to_bf16(q6) == ref).all()
    # extremes: all-ones payload (q = 31, scales = -1), zero payload (q = -32), d = largest finite half
    b = np.zeros((3, 210), np.uint8)
    b[0, :208] = 0xFF; b[0, 208:] = np.array([0x3C00], "<u2").view(np.uint8)
    b[1, 208:] = np.array([0x7BFF], "<u2").view(np.uint8); b[1, 192:208] = 0x7F
    b[2, :192] = 0xA5; b[2, 192:208] = 0x80; b[2, 208:] = np.array([0x8001], "<u2").view(np.uint8)
    ref = bits16(torch.from_numpy(quants.dequantize(b, GGMLQuantizationType.Q6_K)).to(torch.bfloat16))
    assert (oracle.dequant_q6k_bf16(b) == ref).all() and (coracle.q6k_to_bf16(b) == ref).all()


def test_mixed_quant_golden_file_vs_gguf_py():
    import json
    p = os.path.join(G, "q4km_mix.gguf")
    exp = json.load(open(p + ".expected.json"))
    outs = np.load(p + ".bf16.npz")
    raw = open(p, "rb").read()
    seen = set()
    for t in exp["tensors"]:
        rec = dict(name=t["name"], dtype=t["dtype"], shape=t["shape"], nbytes=t["nbytes"])
        got = oracle.convert_tensor(rec, raw[t["file_offset"]:t["file_offset"] + t["nbytes"]]).view(np.uint16)
        assert (got == outs[t["name"]]).all(), t["name"]
        seen.add(t["dtype"])
    assert {"Q4_K", "Q6_K", "Q8_0", "F32"} <= seen


F4_TYPES = ["Q4_0", "Q4_1", "Q5_0", "Q5_1", "Q2_K", "Q3_K", "Q5_K"]


CB_TYPES = ["IQ4_NL", "IQ4_XS", "MXFP4"]
IQ_TYPES = ["IQ2_XXS", "IQ2_XS", "IQ2_S", "IQ3_XXS", "IQ3_S", "IQ1_S", "IQ1_M", "TQ1_0", "TQ2_0", "NVFP4"]
SCALE_OFFS = {"Q4_0": [0], "Q4_1": [0, 2], "Q5_0": [0], "Q5_1": [0, 2], "Q2_K": [80, 82], "Q3_K": [108], "Q5_K": [0, 2], "Q4_K": [0, 2], "Q6_K": [208], "Q8_0": [0],
              "IQ4_NL": [0], "IQ4_XS": [0], "MXFP4": [], "IQ2_XXS": [0], "IQ2_XS": [0], "IQ2_S": [0], "IQ3_XXS": [0], "IQ3_S": [0], "IQ1_S": [0], "IQ1_M": [],
              "TQ1_0": [52], "TQ2_0": [64], "NVFP4": []}


def test_generated_codebooks_equal_gguf_py():
    """oracle/iq_grids.py, oracle/iq_grids_c.h and kukeon_b200/csrc/kk_iq_grids.h are generated files: re-derive them from gguf-py here."""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_iq_grids", os.path.join(root, "tools", "gen_iq_grids.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    from oracle.iq_grids import GRIDS
    T = gen.tables()
    assert set(T) == set(GRIDS)
    for name, g in T.items():
        assert [bytes(r.tolist()) for r in g] == GRIDS[name], name
    h1 = open(os.path.join(root, "kukeon_b200", "csrc", "kk_iq_grids.h")).read()
    assert h1 == open(os.path.join(root, "oracle", "iq_grids_c.h")).read()
    for name, g in T.items():
        w = g.shape[1]
        words = [int.from_bytes(bytes(r.tolist()), "little") for r in g]
        for x in (words[0], words[len(words) // 2], words[-1]):
            assert (f"0x{x:016x}ull" if w == 8 else f"0x{x:08x}u") in h1
    assert (oracle.KSIGNS == np.frombuffer(__import__("gguf").quants.IQ2_XXS.ksigns, np.uint8)).all()


@pytest.mark.parametrize("dtype", F4_TYPES + CB_TYPES + IQ_TYPES + ["Q4_K", "Q6_K", "Q8_0"])
def test_every_block_quant_vs_gguf_py_live_bit_exact(coracle, dtype):
    """§8(f4): numpy restatement and C twin against gguf.quants.dequantize on (a) fully random bytes — Inf/NaN scales
    included, compared as fp32 bit patterns with NaN == NaN — and (b) finite-scale synthetic blocks after RNE to bf16."""
    from gguf import GGMLQuantizationType, quants
    from tools import synth
    nel, nb, fn = oracle.BLOCK_QUANTS[dtype]
    qt = getattr(GGMLQuantizationType, dtype)
    rng = np.random.Generator(np.random.Philox(key=1234))
    b = rng.integers(0, 256, size=(4096, nb), dtype=np.uint8)
    with np.errstate(all="ignore"):
        ref, got = quants.dequantize(b, qt), fn(b)
    assert ((ref.view(np.uint32) == got.view(np.uint32)) | (np.isnan(ref) & np.isnan(got))).all()
    b = synth.gen_bytes(dtype, nb * 5000, 9, 4).reshape(-1, nb)
    ref = bits16(torch.from_numpy(quants.dequantize(b, qt)).to(torch.bfloat16))
    assert (oracle.dequant_bf16(dtype, b) == ref).all() and (coracle.dequant_to_bf16(dtype, b) == ref).all()
    # extremes: all-zero and all-ones payloads with d = dmin = 1.0 and the largest finite half
    for fill in (0x00, 0xFF):
        for h in (0x3C00, 0x7BFF, 0xBC00):
            e = np.full((2, nb), fill, np.uint8)
            for off in SCALE_OFFS[dtype]:
                e[:, off:off + 2] = np.array([h], "<u2").view(np.uint8)
            with np.errstate(all="ignore"):
                ref = bits16(torch.from_numpy(quants.dequantize(e, qt)).to(torch.bfloat16))
            nan = (ref & 0x7FFF) > 0x7F80
            got, cgot = oracle.dequant_bf16(dtype, e), coracle.dequant_to_bf16(dtype, e)
            assert ((got == ref) | (nan & (got == 0x7FFF))).all() and (got == cgot).all()


def test_mxfp4_every_shared_exponent_vs_gguf_py(coracle):
    """All 256 E8M0 bytes (0 and 1 are fp32 subnormals, 255 = 2^127) x all 16 code points."""
    from gguf import GGMLQuantizationType, quants
    b = np.zeros((256, 17), np.uint8)
    b[:, 0] = np.arange(256)
    b[:, 1:9] = (np.arange(8, dtype=np.uint8) * 2) | ((np.arange(8, dtype=np.uint8) * 2 + 1) << 4)  # nibbles 0..15 in the low/high halves
    b[:, 9:17] = b[:, 1:9][:, ::-1]
    with np.errstate(all="ignore"):
        ref = bits16(torch.from_numpy(quants.dequantize(b, GGMLQuantizationType.MXFP4)).to(torch.bfloat16))
    assert (oracle.dequant_bf16("MXFP4", b) == ref).all() and (coracle.dequant_to_bf16("MXFP4", b) == ref).all()


@pytest.mark.parametrize("fixture,types", [("quants_f4.gguf", F4_TYPES), ("quants_cb.gguf", CB_TYPES), ("quants_iq.gguf", IQ_TYPES)])
def test_legacy_k_and_codebook_quant_golden_files_vs_gguf_py(fixture, types):
    import json
    p = os.path.join(G, fixture)
    exp = json.load(open(p + ".expected.json"))
    outs = np.load(p + ".bf16.npz")
    raw = open(p, "rb").read()
    shards, recs = oracle.index_path(p)
    assert [(r["name"], r["dtype"], r["shape"], r["file_offset"], r["nbytes"]) for r in recs] == \
        [(t["name"], t["dtype"], t["shape"], t["file_offset"], t["nbytes"]) for t in exp["tensors"]]
    seen = set()
    for t in exp["tensors"]:
        rec = dict(name=t["name"], dtype=t["dtype"], shape=t["shape"], nbytes=t["nbytes"])
        got = oracle.convert_tensor(rec, raw[t["file_offset"]:t["file_offset"] + t["nbytes"]]).view(np.uint16)
        assert (got == outs[t["name"]]).all(), t["name"]
        seen.add(t["dtype"])
    assert set(types) <= seen


def test_fp8_widening_vs_torch_every_bit_pattern(coracle):
    """KK_LOAD_F8_TO_BF16: all 256 E4M3 ("fn") and E5M2 patterns against torch's float8 -> bfloat16 cast; NaNs -> 0x7FFF."""
    b = np.arange(256, dtype=np.uint8)
    for tdt, name, fn in ((torch.float8_e4m3fn, "F8_E4M3", oracle.f8e4m3_bits_to_bf16), (torch.float8_e5m2, "F8_E5M2", oracle.f8e5m2_bits_to_bf16)):
        ref = bits16(torch.from_numpy(b.copy()).view(tdt).to(torch.bfloat16))
        nan = (ref & 0x7FFF) > 0x7F80
        got = fn(b)
        assert int(nan.sum()) == (2 if name == "F8_E4M3" else 6)
        assert (got[~nan] == ref[~nan]).all() and (got[nan] == 0x7FFF).all()
        assert (coracle.f8_to_bf16(name, b) == got).all()
        # widening is exact: going back to float8 reproduces the byte
        back = torch.from_numpy(got[~nan].view(np.int16)).view(torch.bfloat16).to(tdt).view(torch.uint8).numpy()
        assert (back == b[~nan]).all()
