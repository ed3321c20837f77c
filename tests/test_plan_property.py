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


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
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
