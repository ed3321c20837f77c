"""Host planner (kk_plan_describe, CPU only): pool layout == oracle.plan_pool, and executing the planned
reads + segments with the oracle's arithmetic reproduces oracle.expected_pool bit for bit — for SINGLE,
BROADCAST (parts partition the pool) and SCATTER (per-rank slices), with chunk sizes small enough to force
tensors to be split across chunks."""
import os

import numpy as np
import pytest

from kukeon_b200 import gpupool
from oracle import oracle
from tests import helpers
from tools import synth

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MB = 1 << 20


def check_layout(plan, recs, mode, flags, n_parts):
    n_layouts = n_parts if mode == gpupool.MODE_SCATTER else 1
    assert len(plan["layouts"]) == n_layouts
    for L in range(n_layouts):
        want, total = oracle.plan_pool(recs, mode, flags, n_parts, L)
        lay = plan["layouts"][L]
        assert lay["pool_bytes"] == total
        for w, g in zip(want, lay["tensors"]):
            assert (g["name"], g["dtype"], g["shape"], g["pool_offset"], g["nbytes"]) == \
                   (w["name"], w["dtype"], w["shape"], w["pool_offset"], w["nbytes"])
            assert g["slice_dim"] == w["slice_dim"] and g["slice_begin"] == w["slice_begin"]
            assert g["pool_offset"] % 256 == 0


def run_case(path, mode=gpupool.MODE_SINGLE, flags=0, n_parts=1, chunk=2 * MB):
    shards, recs = oracle.index_path(path)
    plan = gpupool.plan_describe(path, mode=mode, flags=flags, n_parts=n_parts, chunk_bytes=chunk)
    check_layout(plan, recs, mode, flags, n_parts)
    if mode == gpupool.MODE_SCATTER:
        for g in range(n_parts):
            exp, pl = oracle.expected_pool(shards, recs, mode, flags, n_parts, g)
            got, mask = helpers.emulate_part(plan, g, len(exp))
            assert (mask == helpers.expected_mask(pl, len(exp))).all()
            assert (got == exp).all()
    else:
        exp, pl = oracle.expected_pool(shards, recs, mode, flags)
        acc = np.zeros(len(exp), np.uint8)
        cover = np.zeros(len(exp), np.int32)
        for g in range(n_parts):
            got, mask = helpers.emulate_part(plan, g, len(exp))
            acc[mask] = got[mask]
            cover += mask
        assert (cover == helpers.expected_mask(pl, len(exp)).astype(np.int32)).all(), "parts must partition the pool exactly"
        assert (acc == exp).all()
        assert sum(p["src_bytes"] for p in plan["parts"]) == plan["file_bytes"]
    for p in plan["parts"]:
        for ch in p["chunks"]:
            assert ch["buf_bytes"] <= chunk
    return plan


def test_mixed_safetensors_single(native, tmp_path):
    p = str(tmp_path / "m.safetensors")
    helpers.mixed_safetensors(p)
    plan = run_case(p)
    assert len(plan["parts"][0]["chunks"]) >= 1


def test_mixed_unpadded_header_single(native, tmp_path):
    p = str(tmp_path / "m.safetensors")
    helpers.mixed_safetensors(p, pad_header=False)
    run_case(p)


def test_keep_f32_flag(native, tmp_path):
    p = str(tmp_path / "m.safetensors")
    helpers.mixed_safetensors(p)
    run_case(p, flags=gpupool.LOAD_KEEP_F32)


def test_golden_files_plan(native):
    run_case(os.path.join(G, "st_mixed.safetensors"))
    run_case(os.path.join(G, "q4k.gguf"))
    run_case(os.path.join(G, "sharded"))


def test_llama_split_across_chunks_and_broadcast_parts(native, tmp_path):
    d = str(tmp_path / "llama")
    synth.make_llama(d, dict(hidden=256, ffn=704, layers=3, kv_dim=64, vocab=3000), max_shard_bytes=3_000_000)
    plan = run_case(d, chunk=1 * MB)
    assert any(len(ch["segs"]) > 1 for ch in plan["parts"][0]["chunks"])
    for n in (2, 3, 8):
        pb = run_case(d, mode=gpupool.MODE_BROADCAST, n_parts=n, chunk=1 * MB)
        srcs = [p["src_bytes"] for p in pb["parts"]]
        assert sum(srcs) == pb["file_bytes"]
        assert max(srcs) - min(srcs) <= 2 * MB, "parts should be balanced to within ~a chunk"


def test_scatter_slices(native, tmp_path):
    d = str(tmp_path / "llama")
    synth.make_llama(d, dict(hidden=256, ffn=704, layers=2, kv_dim=64, vocab=1024), max_shard_bytes=2_500_000)
    for n in (2, 4, 8):
        plan = run_case(d, mode=gpupool.MODE_SCATTER, n_parts=n, chunk=1 * MB)
        lay = {t["name"]: t for t in plan["layouts"][1]["tensors"]}
        assert lay["model.layers.0.self_attn.q_proj.weight"]["slice_dim"] == 0
        assert lay["model.layers.0.self_attn.o_proj.weight"]["slice_dim"] == 1
        assert lay["model.layers.0.mlp.down_proj.weight"]["shape"] == [256, 704 // n]
        assert lay["model.layers.0.input_layernorm.weight"]["slice_dim"] is None
        assert lay["model.embed_tokens.weight"]["slice_begin"] == 1024 // n
        # each rank reads ~1/n of the sliceable bytes (norms are replicated)
        assert plan["parts"][0]["src_bytes"] < plan["file_bytes"] / n * 1.05 + 64 * 1024


def run_case_scatter_exchange(path, flags, n, chunk=1 * MB):
    """SCATTER with KK_LOAD_SCATTER_EXCHANGE set in `flags`: replay every rank (its ROWSPLIT tiles write into the other ranks' pools)
    and compare every rank's pool with the oracle's slice pool.  The oracle does not know the exchange flag: it changes the
    route, not the result."""
    shards, recs = oracle.index_path(path)
    oflags = flags & ~gpupool.LOAD_SCATTER_EXCHANGE
    plan = gpupool.plan_describe(path, mode=gpupool.MODE_SCATTER, flags=flags, n_parts=n, chunk_bytes=chunk)
    check_layout(plan, recs, gpupool.MODE_SCATTER, oflags, n)
    exps = [oracle.expected_pool(shards, recs, gpupool.MODE_SCATTER, oflags, n, g) for g in range(n)]
    ex = {g: (np.zeros(len(exps[g][0]), np.uint8), np.zeros(len(exps[g][0]), bool)) for g in range(n)}
    for g in range(n):
        got, mask = helpers.emulate_part(plan, g, len(exps[g][0]), exchange=ex)
        assert not (mask & ex[g][1]).any()
        ex[g][0][mask] = got[mask]
        ex[g][1][mask] = True
    for g in range(n):
        exp, pl = exps[g]
        assert (ex[g][1] == helpers.expected_mask(pl, len(exp))).all()
        assert (ex[g][0] == exp).all()
    return plan


def test_scatter_exchange_rows_dealt_to_every_pool(native, tmp_path):
    """KK_LOAD_SCATTER_EXCHANGE: row-parallel tensors are ingested as whole rows by the rank owning 1/N of the rows and
    dealt column-slice by column-slice to all N pools; emulating every rank must reproduce every rank's oracle pool."""
    d = str(tmp_path / "llama")
    synth.make_llama(d, dict(hidden=256, ffn=704, layers=2, kv_dim=64, vocab=1024), max_shard_bytes=2_500_000)
    shards, recs = oracle.index_path(d)
    X = 0x8
    for n in (2, 4, 8):
        plan = gpupool.plan_describe(d, mode=gpupool.MODE_SCATTER, flags=X, n_parts=n, chunk_bytes=1 * MB)
        check_layout(plan, recs, gpupool.MODE_SCATTER, 0, n)
        exps = [oracle.expected_pool(shards, recs, gpupool.MODE_SCATTER, 0, n, g) for g in range(n)]
        ex = {g: (np.zeros(len(exps[g][0]), np.uint8), np.zeros(len(exps[g][0]), bool)) for g in range(n)}
        ops = set()
        for g in range(n):
            got, mask = helpers.emulate_part(plan, g, len(exps[g][0]), exchange=ex)
            assert not (mask & ex[g][1]).any()
            ex[g][0][mask] = got[mask]
            ex[g][1][mask] = True
            ops |= {sg["op"] for ch in plan["parts"][g]["chunks"] for sg in ch["segs"]}
            # whole-row ingest: every read of this rank is one long contiguous range, no 2-7 KB row runs
            n_reads = sum(len(ch["reads"]) for ch in plan["parts"][g]["chunks"])
            assert n_reads <= 2 * len(recs), "exchange plans must not gather rows one pread at a time"
        assert helpers.OP_ROWSPLIT in ops
        for g in range(n):
            exp, pl = exps[g]
            assert (ex[g][1] == helpers.expected_mask(pl, len(exp))).all()
            assert (ex[g][0] == exp).all()
        # every byte of the row-parallel tensors is read exactly once across ranks
        total = sum(p["src_bytes"] for p in plan["parts"])
        repl = sum(r["nbytes"] for r in recs if oracle.slice_dim(r, n) is None)
        assert total == plan["file_bytes"] + (n - 1) * repl


def test_scatter_indivisible_dims_are_replicated(native, tmp_path):
    d = str(tmp_path / "llama")
    synth.make_llama(d, dict(hidden=96, ffn=100, layers=1, kv_dim=32, vocab=77), max_shard_bytes=10_000_000)
    plan = run_case(d, mode=gpupool.MODE_SCATTER, n_parts=8, chunk=1 * MB)
    lay = {t["name"]: t for t in plan["layouts"][3]["tensors"]}
    assert lay["model.embed_tokens.weight"]["slice_dim"] is None  # 77 % 8 != 0
    assert lay["model.layers.0.self_attn.q_proj.weight"]["slice_dim"] == 0  # 96 % 8 == 0


def test_gpt2_conv1d_transpose(native, tmp_path):
    p = str(tmp_path / "gpt2.safetensors")
    synth.make_gpt2(p, n_layer=2, d=96, vocab=301, n_pos=40)
    plan = run_case(p, flags=gpupool.LOAD_GPT2_CONV1D_T, chunk=1 * MB)
    lay = {t["name"]: t for t in plan["layouts"][0]["tensors"]}
    assert lay["h.0.attn.c_attn.weight"]["shape"] == [288, 96]
    assert lay["h.1.mlp.c_proj.weight"]["shape"] == [96, 384]
    assert lay["wte.weight"]["shape"] == [301, 96]
    run_case(p, flags=gpupool.LOAD_GPT2_CONV1D_T | gpupool.LOAD_KEEP_F32, chunk=1 * MB)
    run_case(p, flags=gpupool.LOAD_GPT2_CONV1D_T, mode=gpupool.MODE_BROADCAST, n_parts=2, chunk=1 * MB)


def test_gpt2_conv1d_transpose_shapes_dtypes_and_wide_rows(native, tmp_path):
    """One transpose family (8-row tiles): every dtype, destination rows that are not 16-byte multiples (d = 41, 43: scalar stores), rows wider
    than one tile (3096 and 4128 columns, cut into equal pieces), KEEP_F32 (4-byte outputs), fan-out over 3 parts."""
    T = gpupool.LOAD_GPT2_CONV1D_T
    tr = {helpers.OP_T_F32_BF16, helpers.OP_T_F16_BF16, helpers.OP_T_B16, helpers.OP_T_B32}
    p = str(tmp_path / "gpt2.safetensors")
    synth.make_gpt2(p, n_layer=2, d=96, vocab=301, n_pos=40)
    plan = run_case(p, flags=T, chunk=1 * MB)
    assert {sg["op"] for ch in plan["parts"][0]["chunks"] for sg in ch["segs"]} & tr
    run_case(p, flags=T, mode=gpupool.MODE_BROADCAST, n_parts=3, chunk=1 * MB)
    plan = run_case(p, flags=T | gpupool.LOAD_KEEP_F32, chunk=1 * MB)
    assert {sg["op"] for ch in plan["parts"][0]["chunks"] for sg in ch["segs"]} & {helpers.OP_T_B32}
    for dt, d in (("F16", 40), ("BF16", 40), ("F32", 41), ("F16", 43)):
        q = str(tmp_path / f"gpt2_{dt}_{d}.safetensors")
        synth.write_safetensors(q, synth.gpt2_tensors(n_layer=1, d=d, vocab=50, n_pos=8, dtype=dt), 3)
        plan = run_case(q, flags=T, chunk=1 * MB)
        assert {sg["op"] for ch in plan["parts"][0]["chunks"] for sg in ch["segs"]} & tr, (dt, d)
    # rows wider than one tile (a whole d = 1032 layer would be 50 MB of pools and hit maps; 72 rows of it carry the same tile cuts)
    q = str(tmp_path / "gpt2_wide.safetensors")
    synth.write_safetensors(q, [("h.0.attn.c_attn.weight", "F32", [72, 3096]), ("h.0.mlp.c_fc.weight", "F32", [72, 4128]), ("h.0.mlp.c_proj.weight", "F16", [4128, 72])], 3)
    plan = run_case(q, flags=T, chunk=1 * MB)
    assert {sg["op"] for ch in plan["parts"][0]["chunks"] for sg in ch["segs"]} & tr


def test_gpt2_f16_and_bf16_transpose(native, tmp_path):
    for dt, d in (("F16", 40), ("BF16", 40), ("F32", 41), ("F16", 43)):
        p = str(tmp_path / f"gpt2_{dt}_{d}.safetensors")
        synth.write_safetensors(p, synth.gpt2_tensors(n_layer=1, d=d, vocab=50, n_pos=8, dtype=dt), 3)
        run_case(p, flags=gpupool.LOAD_GPT2_CONV1D_T, chunk=1 * MB)


def test_mixtral_gguf_plan(native, tmp_path):
    p = str(tmp_path / "mix.gguf")
    synth.write_gguf(p, synth.mixtral_gguf_tensors(hidden=256, ffn=768, layers=2, experts=2, vocab=512, kv_dim=256), 7)
    run_case(p, chunk=1 * MB)
    run_case(p, mode=gpupool.MODE_BROADCAST, n_parts=4, chunk=1 * MB)
    run_case(p, mode=gpupool.MODE_SCATTER, n_parts=2, chunk=1 * MB)


def test_unsupported_quant_is_refused_at_plan_time(native, tmp_path):
    """Q8_K is llama.cpp's intermediate activation format (never a weight type in released files) and the one ggml block
    type the index knows but the kernel does not dequantise."""
    import struct
    p = str(tmp_path / "q8k.gguf")
    head = struct.pack("<IIQQ", 0x46554747, 3, 1, 0) + struct.pack("<Q", 1) + b"w" + struct.pack("<I", 1) + struct.pack("<Q", 256) + struct.pack("<IQ", 15, 0)
    head += b"\0" * ((-len(head)) % 32)
    open(p, "wb").write(head + b"\0" * 320)
    assert gpupool.index(p)[0]["dtype"] == "Q8_K"  # indexing works ...
    with pytest.raises(gpupool.ErrUnsupported, match="Q8_K"):  # ... loading is refused, never approximated
        gpupool.plan_describe(p)


F4_MIX = ["Q4_0", "Q4_1", "Q5_0", "Q5_1", "Q2_K", "Q3_K", "Q5_K", "IQ4_NL", "IQ4_XS", "MXFP4", "IQ2_XXS", "IQ2_XS", "IQ2_S", "IQ3_XXS", "IQ3_S", "IQ1_S",
          "IQ1_M", "TQ1_0", "TQ2_0", "NVFP4"]


def f4_tensors(hidden=256, ffn=768, layers=2, vocab=512):
    """Every §8(f4) type in one llama-shaped inventory (column-parallel names so that SCATTER slices whole rows of blocks)."""
    t = [("token_embd.weight", "Q2_K", [vocab, hidden])]
    for i in range(layers):
        p = f"blk.{i}."
        t += [(p + "attn_norm.weight", "F32", [hidden]), (p + "attn_q.weight", "Q4_0", [hidden, hidden]), (p + "attn_k.weight", "Q4_1", [hidden // 4, hidden]),
              (p + "attn_v.weight", "Q5_0", [hidden // 4, hidden]), (p + "attn_output.weight", "Q5_1", [hidden, hidden]),
              (p + "ffn_gate.weight", "Q3_K", [ffn, hidden]), (p + "ffn_up.weight", "Q5_K", [ffn, hidden]), (p + "ffn_down.weight", "Q2_K", [hidden, ffn]),
              (p + "ffn_gate_exps.weight", "MXFP4", [2, ffn, hidden]), (p + "ffn_up_exps.weight", "IQ4_XS", [2, ffn, hidden]),
              (p + "ffn_down_exps.weight", "IQ4_NL", [2, hidden, ffn]),
              (p + "a.weight", "IQ2_XXS", [hidden, hidden]), (p + "b.weight", "IQ2_XS", [hidden // 4, hidden]), (p + "c.weight", "IQ2_S", [hidden // 4, hidden]),
              (p + "d.weight", "IQ3_XXS", [hidden, hidden]), (p + "e.weight", "IQ3_S", [ffn, hidden]), (p + "f.weight", "IQ1_S", [ffn, hidden]),
              (p + "g.weight", "IQ1_M", [hidden, ffn]), (p + "h.weight", "TQ1_0", [hidden, hidden]), (p + "i.weight", "TQ2_0", [hidden, hidden]),
              (p + "j.weight", "NVFP4", [hidden, hidden])]
    t += [("output_norm.weight", "F32", [hidden]), ("output.weight", "Q5_K", [vocab, hidden])]
    return t


def test_legacy_and_k_quant_plan(native, tmp_path):
    p = str(tmp_path / "f4.gguf")
    synth.write_gguf(p, f4_tensors(), 11)
    plan = run_case(p, chunk=1 * MB)
    ops = {sg["op"] for ch in plan["parts"][0]["chunks"] for sg in ch["segs"]}
    assert {helpers.OP_Q4_0, helpers.OP_Q4_1, helpers.OP_Q5_0, helpers.OP_Q5_1, helpers.OP_Q2K, helpers.OP_Q3K, helpers.OP_Q5K, helpers.OP_IQ4NL,
            helpers.OP_IQ4XS, helpers.OP_MXFP4, helpers.OP_IQ2XXS, helpers.OP_IQ2XS, helpers.OP_IQ2S, helpers.OP_IQ3XXS, helpers.OP_IQ3S, helpers.OP_IQ1S,
            helpers.OP_IQ1M, helpers.OP_TQ1_0, helpers.OP_TQ2_0, helpers.OP_NVFP4} <= ops
    run_case(p, mode=gpupool.MODE_BROADCAST, n_parts=3, chunk=1 * MB)
    plan = run_case(p, mode=gpupool.MODE_SCATTER, n_parts=4, chunk=1 * MB)
    lay = {t["name"]: t for t in plan["layouts"][1]["tensors"]}
    assert lay["blk.0.ffn_up.weight"]["slice_dim"] == 0 and lay["blk.0.ffn_up.weight"]["shape"] == [192, 256]
    assert lay["blk.0.ffn_down.weight"]["slice_dim"] is None  # dim-1 slices would cut quantised blocks
    run_case(os.path.join(G, "quants_f4.gguf"))
    run_case(os.path.join(G, "quants_cb.gguf"))
    run_case(os.path.join(G, "quants_iq.gguf"))
    # tensors bigger than one tile of every type, so that tile boundaries inside a tensor are exercised (Q4_0: 1816 blocks/tile)
    big = [(f"blk.{i}.ffn_up.weight", dt, [96, 2048]) for i, dt in enumerate(F4_MIX)]
    p2 = str(tmp_path / "f4big.gguf")
    synth.write_gguf(p2, big, 12)
    run_case(p2, chunk=1 * MB)


def q4km_tensors(hidden=256, ffn=768, layers=2, vocab=512):
    """llama.cpp Q4_K_M-style mix: attention/ffn in Q4_K, ffn_down and output in Q6_K, one Q8_0 tensor, norms F32."""
    t = [("token_embd.weight", "Q4_K", [vocab, hidden])]
    for i in range(layers):
        p = f"blk.{i}."
        t += [(p + "attn_norm.weight", "F32", [hidden]), (p + "attn_q.weight", "Q4_K", [hidden, hidden]),
              (p + "attn_v.weight", "Q8_0", [hidden // 4, hidden]), (p + "ffn_up.weight", "Q4_K", [ffn, hidden]),
              (p + "ffn_down.weight", "Q6_K", [hidden, ffn])]
    t += [("output_norm.weight", "F32", [hidden]), ("output.weight", "Q6_K", [vocab, hidden])]
    return t


def test_q4_k_m_style_mixed_quant_plan(native, tmp_path):
    p = str(tmp_path / "q4km.gguf")
    synth.write_gguf(p, q4km_tensors(), 9)
    run_case(p, chunk=1 * MB)
    run_case(p, mode=gpupool.MODE_BROADCAST, n_parts=3, chunk=1 * MB)
    run_case(p, mode=gpupool.MODE_SCATTER, n_parts=2, chunk=1 * MB)
    run_case(os.path.join(G, "q4km_mix.gguf"))


def test_fp8_checkpoint_verbatim_by_default_widened_on_request(native, tmp_path):
    """FP8 safetensors (DeepSeek-V3 / Llama-3.1-FP8 style: F8_E4M3 weights + F32 per-block scale tensors): verbatim bytes by
    default, bf16 with KK_LOAD_F8_TO_BF16; dim-1 scatter slices stay element-granular."""
    p = str(tmp_path / "fp8.safetensors")
    t = [("model.embed_tokens.weight", "BF16", [512, 128])]
    for i in range(2):
        q = f"model.layers.{i}."
        t += [(q + "self_attn.q_proj.weight", "F8_E4M3", [128, 128]), (q + "self_attn.q_proj.weight_scale_inv", "F32", [1, 1]),
              (q + "self_attn.o_proj.weight", "F8_E4M3", [128, 128]), (q + "mlp.down_proj.weight", "F8_E5M2", [128, 352]),
              (q + "mlp.up_proj.weight", "F8_E4M3", [352, 128]), (q + "input_layernorm.weight", "BF16", [128]), (q + "odd.weight", "F8_E4M3", [3, 37])]
    synth.write_safetensors(p, t, 17)
    plan = run_case(p, chunk=1 * MB)
    assert {x["dtype"] for x in plan["layouts"][0]["tensors"]} >= {"F8_E4M3", "F8_E5M2"}
    plan = run_case(p, flags=gpupool.LOAD_F8_TO_BF16, chunk=1 * MB)
    assert not any(x["dtype"].startswith("F8") for x in plan["layouts"][0]["tensors"])
    ops = {sg["op"] for ch in plan["parts"][0]["chunks"] for sg in ch["segs"]}
    assert {helpers.OP_F8E4M3, helpers.OP_F8E5M2} <= ops
    run_case(p, flags=gpupool.LOAD_F8_TO_BF16, mode=gpupool.MODE_BROADCAST, n_parts=3, chunk=1 * MB)
    for flags in (0, gpupool.LOAD_F8_TO_BF16, gpupool.LOAD_F8_TO_BF16 | gpupool.LOAD_SCATTER_EXCHANGE):
        plan = run_case_scatter_exchange(p, flags, 4) if flags & gpupool.LOAD_SCATTER_EXCHANGE else run_case(p, mode=gpupool.MODE_SCATTER, flags=flags, n_parts=4, chunk=1 * MB)
        lay = {x["name"]: x for x in plan["layouts"][1]["tensors"]}
        assert lay["model.layers.0.mlp.down_proj.weight"]["slice_dim"] == 1 and lay["model.layers.0.mlp.down_proj.weight"]["shape"] == [128, 88]


def test_bad_arguments(native, tmp_path):
    p = str(tmp_path / "m.safetensors")
    helpers.mixed_safetensors(p)
    with pytest.raises(gpupool.ErrInvalid):
        gpupool.plan_describe(p, mode=gpupool.MODE_SINGLE, n_parts=2)
    with pytest.raises(gpupool.ErrInvalid):
        gpupool.plan_describe(p, mode=7)
    with pytest.raises(gpupool.ErrInvalid):
        gpupool.plan_describe(p, mode=gpupool.MODE_BROADCAST, n_parts=9)


def test_scatter_of_empty_2d_tensors_does_not_divide_by_zero(native, tmp_path):
    """Found by tools/fuzz (UBSan): a [0, C] tensor with a column-parallel name used to divide by its row count."""
    p = str(tmp_path / "empty.safetensors")
    hdr = {"model.layers.0.self_attn.q_proj.weight": {"dtype": "BF16", "shape": [0, 64], "data_offsets": [0, 0]},
           "model.layers.0.self_attn.o_proj.weight": {"dtype": "BF16", "shape": [8, 0], "data_offsets": [0, 0]},
           "model.layers.0.mlp.down_proj.weight": {"dtype": "BF16", "shape": [8, 16], "data_offsets": [0, 256]}}
    helpers.write_raw_safetensors(p, hdr, bytes(range(256)))
    for flags in (0, 0x8):
        plan = gpupool.plan_describe(p, mode=gpupool.MODE_SCATTER, flags=flags, n_parts=4, chunk_bytes=1 * MB)
        lay = {t["name"]: t for t in plan["layouts"][2]["tensors"]}
        assert lay["model.layers.0.self_attn.q_proj.weight"]["slice_dim"] is None and lay["model.layers.0.self_attn.q_proj.weight"]["nbytes"] == 0
        assert lay["model.layers.0.mlp.down_proj.weight"]["slice_dim"] == 1
    run_case(p, mode=gpupool.MODE_SCATTER, n_parts=4, chunk=1 * MB)


def test_scatter_never_slices_sub_byte_dtypes(native, tmp_path):
    """an F4 q_proj.weight of shape [4, 3] (6 bytes) in SCATTER with n = 2 was sliced along dim 0 as if rows were byte aligned (two
    2-byte slices at +0 and +2 instead of 3-byte rows): the sub-byte check sat in a dead branch.  F4 / F6 tensors are replicated whole."""
    p = str(tmp_path / "f4.safetensors")
    hdr = {"model.layers.0.self_attn.q_proj.weight": {"dtype": "F4", "shape": [4, 3], "data_offsets": [0, 6]},
           "model.layers.0.self_attn.k_proj.weight": {"dtype": "F6_E2M3", "shape": [4, 4], "data_offsets": [6, 18]},
           "model.layers.0.self_attn.v_proj.weight": {"dtype": "BF16", "shape": [4, 4], "data_offsets": [18, 50]}}
    helpers.write_raw_safetensors(p, hdr, bytes(range(50)))
    plan = gpupool.plan_describe(p, mode=gpupool.MODE_SCATTER, n_parts=2, chunk_bytes=1 * MB)
    for g in range(2):
        lay = {t["name"]: t for t in plan["layouts"][g]["tensors"]}
        assert lay["model.layers.0.self_attn.q_proj.weight"]["slice_dim"] is None and lay["model.layers.0.self_attn.q_proj.weight"]["nbytes"] == 6
        assert lay["model.layers.0.self_attn.k_proj.weight"]["slice_dim"] is None and lay["model.layers.0.self_attn.k_proj.weight"]["nbytes"] == 12
        assert lay["model.layers.0.self_attn.v_proj.weight"]["slice_dim"] == 0 and lay["model.layers.0.self_attn.v_proj.weight"]["nbytes"] == 16
    run_case(p, mode=gpupool.MODE_SCATTER, n_parts=2, chunk=1 * MB)
