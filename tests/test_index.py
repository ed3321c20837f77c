"""kk_index (product, C++) == oracle.index_path (Python restatement) == the format owners' readers,
on the committed golden files, on synthetic files, and on the malformed-header cases safetensors rejects
(SURVEY.md Appendix C.3)."""
import json
import os
import struct

import numpy as np
import pytest

from kukeon_b200 import gpupool
from oracle import oracle
from tests import helpers
from tools import synth

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def both(path):
    a = gpupool.index(path)
    shards, b = oracle.index_path(path)
    assert a == b, "product index and oracle index must be identical records"
    assert gpupool.index_shards(path) == shards
    return a, shards


def test_golden_safetensors_vs_library(native):
    import hashlib
    from safetensors import safe_open
    p = os.path.join(G, "st_mixed.safetensors")
    recs, _ = both(p)
    exp = json.load(open(p + ".expected.json"))["tensors"]
    by = {r["name"]: r for r in recs}
    assert sorted(by) == sorted(e["name"] for e in exp)
    raw = open(p, "rb").read()
    for e in exp:
        r = by[e["name"]]
        assert r["dtype"] == e["dtype"] and r["shape"] == e["shape"] and r["nbytes"] == e["nbytes"]
        assert hashlib.sha256(raw[r["file_offset"]:r["file_offset"] + r["nbytes"]]).hexdigest() == e["sha256"]
    # live cross-check against the installed reader as well
    with safe_open(p, "pt") as f:
        assert sorted(f.keys()) == sorted(by)
    offs = [r["file_offset"] for r in recs]
    assert offs == sorted(offs)


def test_golden_gguf_vs_gguf_py(native):
    p = os.path.join(G, "q4k.gguf")
    recs, _ = both(p)
    exp = json.load(open(p + ".expected.json"))
    want = sorted(exp["tensors"], key=lambda t: t["file_offset"])
    got = [{k: r[k] for k in ("name", "dtype", "shape", "file_offset", "nbytes")} for r in recs]
    assert got == want
    assert all(r["file_offset"] % exp["alignment"] == 0 for r in recs)
    assert min(r["file_offset"] for r in recs) == exp["data_offset"]


def test_golden_sharded_index_json(native):
    import hashlib
    d = os.path.join(G, "sharded")
    recs, shards = both(d)
    wm = json.load(open(os.path.join(d, "model.safetensors.index.json")))["weight_map"]
    exp = json.load(open(os.path.join(d, "expected.json")))
    assert len(shards) == len(set(wm.values())) == 4
    assert [os.path.basename(s) for s in shards] == sorted(set(wm.values()))
    for r in recs:
        assert os.path.basename(shards[r["shard"]]) == wm[r["name"]]
        raw = open(shards[r["shard"]], "rb").read()
        assert hashlib.sha256(raw[r["file_offset"]:r["file_offset"] + r["nbytes"]]).hexdigest() == exp[r["name"]]["sha256"]
    keys = [(r["shard"], r["file_offset"]) for r in recs]
    assert keys == sorted(keys)
    # the index.json path itself and the directory resolve identically
    assert gpupool.index(os.path.join(d, "model.safetensors.index.json")) == recs


def test_synthetic_llama_inventory(native, tmp_path):
    cfg = dict(hidden=64, ffn=160, layers=3, kv_dim=32, vocab=500)
    synth.make_llama(str(tmp_path / "m"), cfg, max_shard_bytes=150_000)
    recs, shards = both(str(tmp_path / "m"))
    assert len(recs) == 9 * 3 + 3 and len(shards) > 1
    assert sum(r["nbytes"] for r in recs) == synth.total_bytes(synth.llama_tensors(**cfg))
    from safetensors import safe_open
    for i, s in enumerate(shards):
        with safe_open(s, "np") as f:
            for r in (r for r in recs if r["shard"] == i):
                sl = f.get_slice(r["name"])
                assert list(sl.get_shape()) == r["shape"] and sl.get_dtype() == r["dtype"]


def test_full_size_inventories_match_survey():
    # tensor counts / byte totals of the BASELINE configs without writing files
    assert len(synth.llama_tensors(**synth.LLAMA3_8B)) == 291
    assert synth.total_bytes(synth.llama_tensors(**synth.LLAMA3_8B)) == 16_060_522_496
    assert len(synth.llama_tensors(**synth.LLAMA3_70B)) == 723
    assert synth.total_bytes(synth.llama_tensors(**synth.LLAMA3_70B)) == 141_107_412_992
    assert len(synth.gpt2_tensors()) == 148
    assert synth.total_bytes(synth.gpt2_tensors()) == 497_759_232
    mt = synth.mixtral_gguf_tensors()
    assert len(mt) == 323
    w = sum(int(np.prod(s)) for _, _, s in mt)
    assert w == 46_702_792_704


def test_synthetic_gguf_vs_gguf_reader(native, tmp_path):
    import gguf
    p = str(tmp_path / "m.gguf")
    synth.write_gguf(p, synth.mixtral_gguf_tensors(hidden=256, ffn=512, layers=2, experts=2, vocab=512, kv_dim=256), 7)
    recs, _ = both(p)
    r = gguf.GGUFReader(p)
    assert len(r.tensors) == len(recs) == 23
    by = {x["name"]: x for x in recs}
    for t in r.tensors:
        x = by[t.name]
        assert (x["file_offset"], x["nbytes"], x["dtype"]) == (int(t.data_offset), int(t.n_bytes), t.tensor_type.name)
        assert x["shape"] == [int(v) for v in reversed(t.shape.tolist())]


def test_unpadded_header_is_accepted(native, tmp_path):
    p = str(tmp_path / "u.safetensors")
    helpers.mixed_safetensors(p, pad_header=False)
    recs, _ = both(p)
    assert len(recs) == 11
    assert {r["name"] for r in recs if r["nbytes"] == 0} == {"f.empty"}


def _hdr(**tensors):
    return {k: {"dtype": v[0], "shape": v[1], "data_offsets": v[2]} for k, v in tensors.items()}


BAD = {
    "hole": (_hdr(a=("F32", [2], [0, 8]), b=("F32", [2], [12, 20])), 20, "invalid offset"),
    "overlap": (_hdr(a=("F32", [2], [0, 8]), b=("F32", [2], [4, 12])), 12, "invalid offset"),
    "trailing": (_hdr(a=("F32", [2], [0, 8])), 12, "not fully covered"),
    "truncated": (_hdr(a=("F32", [4], [0, 16])), 8, "not fully covered"),
    "size_mismatch": (_hdr(a=("F32", [3], [0, 8])), 8, "invalid shape"),
    "bad_dtype": (_hdr(a=("F33", [2], [0, 8])), 8, "unknown variant"),
    "not_json": (b"{\"a\": nope}", 0, "header"),
    "not_object": (b"[1,2,3]", 0, "header"),
}


@pytest.mark.parametrize("case", sorted(BAD))
def test_malformed_safetensors_rejected_like_the_library(native, tmp_path, case):
    from safetensors import safe_open
    hdr, ndata, frag = BAD[case]
    p = str(tmp_path / f"{case}.safetensors")
    helpers.write_raw_safetensors(p, hdr, b"\0" * ndata)
    with pytest.raises(Exception):  # the format owner rejects it ...
        with safe_open(p, "np") as f:
            list(f.keys())
    with pytest.raises(gpupool.ErrFormat) as ei:  # ... and so does the product
        gpupool.index(p)
    assert frag in str(ei.value)
    with pytest.raises(oracle.OracleError):  # ... and the oracle
        oracle.index_safetensors(p)


def test_header_too_large_and_too_small(native, tmp_path):
    p = str(tmp_path / "big.safetensors")
    helpers.write_raw_safetensors(p, b"{}", b"", n_override=200_000_000)
    with pytest.raises(gpupool.ErrFormat, match="header too large"):
        gpupool.index(p)
    q = str(tmp_path / "tiny.safetensors")
    open(q, "wb").write(b"\1\2\3")
    with pytest.raises(gpupool.ErrFormat, match="header too small"):
        gpupool.index(q)
    z = str(tmp_path / "len.safetensors")
    helpers.write_raw_safetensors(z, b"{}", b"", n_override=1000)
    with pytest.raises(gpupool.ErrFormat, match="invalid header length"):
        gpupool.index(z)


def test_zero_tensors_file_ok(native, tmp_path):
    p = str(tmp_path / "none.safetensors")
    helpers.write_raw_safetensors(p, {"__metadata__": {"k": "v"}}, b"")
    assert gpupool.index(p) == []


def test_gguf_errors(native, tmp_path):
    p = str(tmp_path / "bad.gguf")
    open(p, "wb").write(b"GGUX" + b"\0" * 40)
    with pytest.raises(gpupool.ErrFormat, match="magic"):
        gpupool.index(p)
    open(p, "wb").write(struct.pack("<IIQQ", 0x46554747, 3, 5, 0))  # claims 5 tensors, has none
    with pytest.raises(gpupool.ErrFormat, match="truncated"):
        gpupool.index(p)
    open(p, "wb").write(struct.pack("<IIQQ", 0x46554747, 9, 0, 0))
    with pytest.raises(gpupool.ErrUnsupported, match="version"):
        gpupool.index(p)
    

def test_gguf_custom_alignment_and_bad_block_row(native, tmp_path):
    p = str(tmp_path / "a64.gguf")
    synth.write_gguf(p, [("x.weight", "Q4_K", [2, 256]), ("y.weight", "F32", [3])], 1, alignment=64)
    recs, _ = both(p)
    assert all(r["file_offset"] % 64 == 0 for r in recs)
    bad = str(tmp_path / "row.gguf")
    head = struct.pack("<IIQQ", 0x46554747, 3, 1, 0)
    name = b"w"
    head += struct.pack("<Q", 1) + name + struct.pack("<I", 2) + struct.pack("<2Q", 100, 4) + struct.pack("<IQ", 12, 0)
    open(bad, "wb").write(head + b"\0" * 4096)
    with pytest.raises(gpupool.ErrFormat, match="multiple"):
        gpupool.index(bad)


def test_missing_path_and_empty_dir(native, tmp_path):
    with pytest.raises(gpupool.ErrNotFound):
        gpupool.index(str(tmp_path / "nope"))
    os.makedirs(tmp_path / "empty")
    with pytest.raises(gpupool.ErrNotFound):
        gpupool.index(str(tmp_path / "empty"))


def test_weight_map_inconsistency_rejected(native, tmp_path):
    d = tmp_path / "m"
    synth.make_llama(str(d), dict(hidden=32, ffn=64, layers=1, kv_dim=16, vocab=64), max_shard_bytes=9000)
    idx = d / "model.safetensors.index.json"
    doc = json.load(open(idx))
    k = sorted(doc["weight_map"])[0]
    others = sorted(set(doc["weight_map"].values()) - {doc["weight_map"][k]})
    doc["weight_map"][k] = others[0]
    json.dump(doc, open(idx, "w"))
    with pytest.raises(gpupool.ErrFormat, match="weight_map"):
        gpupool.index(str(d))


def test_unicode_and_escaped_tensor_names(native, tmp_path):
    """The header is JSON: names may carry escapes and non-ASCII text; product, oracle and safetensors must agree."""
    from safetensors import safe_open
    names = ["plain", "with space", "quote\"d", "back\\slash", "tab\there", "unicode-é-漢字-😀", "slash/inside", "escé"]
    hdr, off = {}, 0
    for i, n in enumerate(names):
        hdr[n] = {"dtype": "U8", "shape": [i + 1], "data_offsets": [off, off + i + 1]}
        off += i + 1
    p = str(tmp_path / "names.safetensors")
    # ensure_ascii=True makes json emit \\uXXXX escapes (incl. a surrogate pair for the emoji): exercises the parser
    raw = json.dumps(hdr, ensure_ascii=True).encode()
    helpers.write_raw_safetensors(p, raw, bytes(off))
    recs, _ = both(p)
    assert sorted(r["name"] for r in recs) == sorted(names)
    with safe_open(p, "np") as f:
        assert sorted(f.keys()) == sorted(names)
    q = str(tmp_path / "names_utf8.safetensors")
    helpers.write_raw_safetensors(q, json.dumps(hdr, ensure_ascii=False).encode("utf-8"), bytes(off))
    assert [r["name"] for r in gpupool.index(q)] == [r["name"] for r in recs]


def test_limits_name_length_and_rank(native, tmp_path):
    p = str(tmp_path / "long.safetensors")
    helpers.write_raw_safetensors(p, {"n" * 255: {"dtype": "U8", "shape": [1], "data_offsets": [0, 1]}}, b"\0")
    assert len(gpupool.index(p)[0]["name"]) == 255
    helpers.write_raw_safetensors(p, {"n" * 256: {"dtype": "U8", "shape": [1], "data_offsets": [0, 1]}}, b"\0")
    with pytest.raises(gpupool.ErrUnsupported, match="name longer"):
        gpupool.index(p)
    helpers.write_raw_safetensors(p, {"t": {"dtype": "U8", "shape": [1] * 8, "data_offsets": [0, 1]}}, b"\0")
    assert gpupool.index(p)[0]["shape"] == [1] * 8
    helpers.write_raw_safetensors(p, {"t": {"dtype": "U8", "shape": [1] * 9, "data_offsets": [0, 1]}}, b"\0")
    with pytest.raises(gpupool.ErrUnsupported, match="dims"):
        gpupool.index(p)


def test_shape_overflow_and_negative_numbers_rejected(native, tmp_path):
    p = str(tmp_path / "ovf.safetensors")
    helpers.write_raw_safetensors(p, {"t": {"dtype": "F32", "shape": [2 ** 40, 2 ** 40], "data_offsets": [0, 4]}}, b"\0" * 4)
    with pytest.raises(gpupool.ErrFormat):
        gpupool.index(p)
    helpers.write_raw_safetensors(p, b'{"t":{"dtype":"F32","shape":[-1],"data_offsets":[0,4]}}', b"\0" * 4)
    with pytest.raises(gpupool.ErrFormat):
        gpupool.index(p)
    helpers.write_raw_safetensors(p, b'{"t":{"dtype":"F32","shape":[1.5],"data_offsets":[0,4]}}', b"\0" * 4)
    with pytest.raises(gpupool.ErrFormat):
        gpupool.index(p)


def test_split_gguf_directory(native, tmp_path):
    d = tmp_path / "split"
    os.makedirs(d)
    synth.write_gguf(str(d / "m-00001-of-00002.gguf"), [("a.weight", "Q4_K", [4, 256]), ("b.weight", "F32", [7])], 1)
    synth.write_gguf(str(d / "m-00002-of-00002.gguf"), [("c.weight", "Q6_K", [2, 256]), ("d.weight", "Q8_0", [3, 64])], 2)
    recs, shards = both(str(d))
    assert len(shards) == 2 and [r["shard"] for r in recs] == [0, 0, 1, 1]
    assert {r["name"]: r["dtype"] for r in recs} == {"a.weight": "Q4_K", "b.weight": "F32", "c.weight": "Q6_K", "d.weight": "Q8_0"}
    synth.write_gguf(str(d / "m-00003-of-00003.gguf"), [("a.weight", "F32", [1])], 3)  # duplicate name across shards
    with pytest.raises(gpupool.ErrFormat, match="more than one shard"):
        gpupool.index(str(d))
