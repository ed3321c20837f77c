"""Property test of the planner (CPU only, hypothesis): for random tensor inventories, load modes, rank counts and
staging sizes, executing the planned reads + segments with the oracle's arithmetic reproduces the oracle's pools,
every pool byte is produced exactly once, and nothing reads or writes out of bounds (tests/helpers.emulate_part)."""
import os
import tempfile

import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

from kukeon_b200 import gpupool
from oracle import oracle
from tests import helpers
from tools import synth

SUFFIXES = ["self_attn.q_proj.weight", "self_attn.o_proj.weight", "mlp.down_proj.weight", "mlp.up_proj.weight", "input_layernorm.weight",
            "attn.c_attn.weight", "mlp.c_proj.weight", "misc.weight", "bias"]
ST_DTYPES = ["BF16", "F32", "F16", "U8", "I64", "F8_E4M3", "F8_E5M2"]


@st.composite
def inventories(draw):
    n = draw(st.integers(1, 7))
    out = []
    for i in range(n):
        suf = draw(st.sampled_from(SUFFIXES))
        dt = draw(st.sampled_from(ST_DTYPES))
        nd = draw(st.integers(0, 3))
        shape = [draw(st.sampled_from([0, 1, 2, 3, 8, 16, 24, 40, 64, 96, 130])) for _ in range(nd)]
        out.append((f"model.layers.{i}.{suf}", dt, shape))
    return out


@settings(max_examples=int(os.environ.get("KK_HYP_EXAMPLES", 60)), deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
@given(inv=inventories(), mode=st.sampled_from([0, 1, 2]), n_parts=st.integers(1, 8), flags=st.sampled_from([0, 1, 2, 3, 8, 9, 16, 19, 24, 33, 35, 49, 65, 67]),
       pad=st.booleans(), chunk_mb=st.sampled_from([1, 2]))
def test_random_inventories_plan_to_the_oracle_pool(native, inv, mode, n_parts, flags, pad, chunk_mb):
    if mode == 0:
        n_parts = 1
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "m.safetensors")
        synth.write_safetensors(p, inv, seed=5, pad_header=pad)
        shards, recs = oracle.index_path(p)
        assert gpupool.index(p) == recs
        plan = gpupool.plan_describe(p, mode=mode, flags=flags, n_parts=n_parts, chunk_bytes=chunk_mb << 20)
        oflags = flags & (3 | 16)  # the oracle knows GPT2_CONV1D_T, KEEP_F32 and F8_TO_BF16; SCATTER_EXCHANGE, T8_TILES and TW_TILES change the route, not the result
        if mode == 2:
            exps = [oracle.expected_pool(shards, recs, 2, oflags, n_parts, g) for g in range(n_parts)]
            ex = {g: (np.zeros(len(exps[g][0]), np.uint8), np.zeros(len(exps[g][0]), bool)) for g in range(n_parts)}
            for g in range(n_parts):
                got, mask = helpers.emulate_part(plan, g, len(exps[g][0]), exchange=ex)
                assert not (mask & ex[g][1]).any()
                ex[g][0][mask] = got[mask]
                ex[g][1][mask] = True
            for g in range(n_parts):
                exp, pl = exps[g]
                assert plan["layouts"][g]["pool_bytes"] == len(exp)
                assert (ex[g][1] == helpers.expected_mask(pl, len(exp))).all()
                assert (ex[g][0] == exp).all()
        else:
            exp, pl = oracle.expected_pool(shards, recs, mode, oflags)
            assert plan["layouts"][0]["pool_bytes"] == len(exp)
            acc = np.zeros(len(exp), np.uint8)
            cover = np.zeros(len(exp), np.int32)
            for g in range(n_parts):
                got, mask = helpers.emulate_part(plan, g, len(exp))
                acc[mask] = got[mask]
                cover += mask
            assert (cover == helpers.expected_mask(pl, len(exp)).astype(np.int32)).all()
            assert (acc == exp).all()


GG_TYPES = ["F32", "F16", "BF16", "Q4_0", "Q4_1", "Q5_0", "Q5_1", "Q8_0", "Q2_K", "Q3_K", "Q4_K", "Q5_K", "Q6_K", "IQ4_NL", "IQ4_XS", "MXFP4", "IQ2_XXS",
            "IQ2_XS", "IQ2_S", "IQ3_XXS", "IQ3_S", "IQ1_S", "IQ1_M", "TQ1_0", "TQ2_0", "NVFP4"]
GG_NAMES = ["attn_q.weight", "attn_output.weight", "ffn_up.weight", "ffn_down.weight", "attn_norm.weight", "ffn_gate_exps.weight", "misc.weight"]


@st.composite
def gguf_inventories(draw):
    n = draw(st.integers(1, 6))
    out = []
    for i in range(n):
        dt = draw(st.sampled_from(GG_TYPES))
        blk = synth.GGML[dt][1]
        nd = draw(st.integers(1, 3))
        inner = blk * draw(st.sampled_from([1, 2, 3, 8])) if blk > 1 else draw(st.sampled_from([1, 3, 8, 40, 130]))
        shape = [draw(st.sampled_from([1, 2, 3, 4, 8, 12])) for _ in range(nd - 1)] + [inner]
        out.append((f"blk.{i}.{draw(st.sampled_from(GG_NAMES))}", dt, shape))
    return out


@settings(max_examples=int(os.environ.get("KK_HYP_EXAMPLES", 50)), deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
@given(inv=gguf_inventories(), mode=st.sampled_from([0, 1, 2]), n_parts=st.integers(1, 8), alignment=st.sampled_from([8, 32, 64]), chunk_mb=st.sampled_from([1, 2]))
def test_random_gguf_inventories_plan_to_the_oracle_pool(native, inv, mode, n_parts, alignment, chunk_mb):
    """Every block-quantised type the kernel dequantises, at random shapes / file alignments / modes / rank counts: index == oracle,
    layouts == oracle, and replaying the planned reads + segments with the oracle's arithmetic reproduces the oracle's pools."""
    if mode == 0:
        n_parts = 1
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "m.gguf")
        synth.write_gguf(p, inv, seed=6, alignment=alignment)
        shards, recs = oracle.index_path(p)
        assert gpupool.index(p) == recs
        plan = gpupool.plan_describe(p, mode=mode, n_parts=n_parts, chunk_bytes=chunk_mb << 20)
        if mode == 2:
            for g in range(n_parts):
                exp, pl = oracle.expected_pool(shards, recs, 2, 0, n_parts, g)
                assert plan["layouts"][g]["pool_bytes"] == len(exp)
                got, mask = helpers.emulate_part(plan, g, len(exp))
                assert (mask == helpers.expected_mask(pl, len(exp))).all()
                assert (got == exp).all()
        else:
            exp, pl = oracle.expected_pool(shards, recs, mode, 0)
            assert plan["layouts"][0]["pool_bytes"] == len(exp)
            acc = np.zeros(len(exp), np.uint8)
            cover = np.zeros(len(exp), np.int32)
            for g in range(n_parts):
                got, mask = helpers.emulate_part(plan, g, len(exp))
                acc[mask] = got[mask]
                cover += mask
            assert (cover == helpers.expected_mask(pl, len(exp)).astype(np.int32)).all()
            assert (acc == exp).all()
