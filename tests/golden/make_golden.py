"""Generates the committed golden fixtures with the FORMAT OWNERS' libraries (the reference has no loader and
no fixtures for this path — SURVEY.md §8(c)).  Run once in the build container:

    python tests/golden/make_golden.py

Libraries: safetensors 0.7.0 (writer + reader), gguf 0.19.0 (GGUFWriter, GGUFReader, quants.dequantize),
huggingface_hub save_torch_state_dict (sharded layout), torch 2.11 (fp32/fp16 -> bf16 RNE).
Outputs (all small): st_mixed.safetensors(+.expected.json), q4k.gguf(+.expected.json, .bf16.npy),
sharded/(...), cast_vectors.npz.
"""
import hashlib
import json
import os
import shutil

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def sha(b: bytes) -> str:
    return hashlib.sha256(b).hexdigest()


def bits16(t: torch.Tensor) -> np.ndarray:
    return t.contiguous().view(torch.int16).numpy().view(np.uint16)


def raw_bytes(t: torch.Tensor) -> bytes:
    t = t.contiguous()
    return t.reshape(-1).view(torch.uint8).numpy().tobytes() if t.numel() else b""


def make_safetensors():
    from safetensors import safe_open
    from safetensors.torch import save_file
    g = torch.Generator().manual_seed(20260921)
    tensors = {
        "model.a.weight": torch.randn(17, 33, generator=g) * 0.02,
        "model.b.weight": (torch.randn(9, 130, generator=g) * 3).to(torch.float16),
        "model.c.weight": (torch.randn(65, 64, generator=g)).to(torch.bfloat16),
        "model.ids": torch.arange(11, dtype=torch.int64),
        "model.empty": torch.zeros(0, 4, dtype=torch.float32),
        "model.scalar": torch.tensor(3.25, dtype=torch.float32),
        "model.odd": torch.randn(5, generator=g).to(torch.float16),
    }
    p = os.path.join(HERE, "st_mixed.safetensors")
    save_file(tensors, p, metadata={"format": "pt", "note": "kukeon golden"})
    exp = []
    with safe_open(p, "pt") as f:
        for k in f.keys():
            t = f.get_tensor(k)
            sl = f.get_slice(k)
            e = dict(name=k, dtype=sl.get_dtype(), shape=list(sl.get_shape()), sha256=sha(raw_bytes(t)), nbytes=t.numel() * t.element_size())
            if t.dtype in (torch.float32, torch.float16, torch.bfloat16):
                e["bf16_sha256"] = sha(bits16(t.to(torch.bfloat16)).tobytes())
            exp.append(e)
        meta = f.metadata()
    json.dump(dict(tensors=exp, metadata=meta), open(p + ".expected.json", "w"), indent=1)


def make_gguf():
    import gguf
    from gguf import GGMLQuantizationType as Q
    rng = np.random.Generator(np.random.Philox(key=77))

    def q4k_blocks(rows, cols):
        nb = rows * cols // 256
        b = rng.integers(0, 256, size=(nb, 144), dtype=np.uint8)
        e = rng.integers(5, 12, size=(nb, 2), dtype=np.uint16)
        m = rng.integers(0, 1024, size=(nb, 2), dtype=np.uint16)
        b[:, 0:4] = ((e << 10) | m).astype("<u2").view(np.uint8).reshape(nb, 4)
        return b.reshape(rows, cols // 256 * 144)

    p = os.path.join(HERE, "q4k.gguf")
    w = gguf.GGUFWriter(p, "llama")
    w.add_uint32("llama.block_count", 1)
    w.add_array("tokenizer.ggml.tokens", ["a", "bc", "def"])
    w.add_tensor("blk.0.attn_q.weight", q4k_blocks(4, 512), raw_dtype=Q.Q4_K)
    w.add_tensor("blk.0.ffn_down.weight", q4k_blocks(3, 256), raw_dtype=Q.Q4_K)
    w.add_tensor("blk.0.attn_norm.weight", rng.standard_normal(8).astype(np.float32))
    w.add_tensor("blk.0.misc.weight", rng.standard_normal((5, 3)).astype(np.float16))
    w.write_header_to_file(); w.write_kv_data_to_file(); w.write_tensors_to_file(); w.close()
    r = gguf.GGUFReader(p)
    exp = []
    outs = {}
    for t in r.tensors:
        shape = [int(x) for x in reversed(t.shape.tolist())]
        exp.append(dict(name=t.name, dtype=t.tensor_type.name, shape=shape, file_offset=int(t.data_offset), nbytes=int(t.n_bytes)))
        f32 = gguf.quants.dequantize(np.array(t.data), t.tensor_type)
        outs[t.name] = bits16(torch.from_numpy(np.ascontiguousarray(f32, dtype=np.float32)).to(torch.bfloat16)).reshape(-1)
    json.dump(dict(tensors=exp, alignment=int(r.alignment), data_offset=int(r.data_offset)), open(p + ".expected.json", "w"), indent=1)
    np.savez_compressed(p + ".bf16.npz", **outs)


def make_gguf_mixed_quants():
    """A Q4_K_M-style file: Q4_K + Q6_K + Q8_0 + F32 tensors (real llama.cpp Q4_K_M checkpoints mix Q4_K and Q6_K)."""
    import gguf
    from gguf import GGMLQuantizationType as Q
    rng = np.random.Generator(np.random.Philox(key=78))

    def blocks(nblk, bsz, doffs):
        b = rng.integers(0, 256, size=(nblk, bsz), dtype=np.uint8)
        for off in doffs:
            e = rng.integers(5, 12, size=nblk, dtype=np.uint16)
            m = rng.integers(0, 1024, size=nblk, dtype=np.uint16)
            sgn = rng.integers(0, 2, size=nblk, dtype=np.uint16) << 15
            b[:, off:off + 2] = ((e << 10) | m | sgn).astype("<u2").view(np.uint8).reshape(nblk, 2)
        return b

    p = os.path.join(HERE, "q4km_mix.gguf")
    w = gguf.GGUFWriter(p, "llama")
    w.add_tensor("blk.0.ffn_down.weight", blocks(3 * 2, 210, [208]).reshape(3, 2 * 210), raw_dtype=Q.Q6_K)   # [3, 512]
    w.add_tensor("blk.0.attn_v.weight", blocks(5 * 3, 34, [0]).reshape(5, 3 * 34), raw_dtype=Q.Q8_0)          # [5, 96]
    w.add_tensor("blk.0.attn_q.weight", blocks(2 * 1, 144, [0, 2]).reshape(2, 144), raw_dtype=Q.Q4_K)          # [2, 256]
    w.add_tensor("output.weight", blocks(4 * 1, 210, [208]).reshape(4, 210), raw_dtype=Q.Q6_K)                  # [4, 256]
    w.add_tensor("blk.0.attn_norm.weight", rng.standard_normal(16).astype(np.float32))
    w.write_header_to_file(); w.write_kv_data_to_file(); w.write_tensors_to_file(); w.close()
    r = gguf.GGUFReader(p)
    exp, outs = [], {}
    for t in r.tensors:
        shape = [int(x) for x in reversed(t.shape.tolist())]
        exp.append(dict(name=t.name, dtype=t.tensor_type.name, shape=shape, file_offset=int(t.data_offset), nbytes=int(t.n_bytes)))
        f32 = gguf.quants.dequantize(np.array(t.data), t.tensor_type)
        outs[t.name] = bits16(torch.from_numpy(np.ascontiguousarray(f32, dtype=np.float32)).to(torch.bfloat16)).reshape(-1)
    json.dump(dict(tensors=exp, alignment=int(r.alignment), data_offset=int(r.data_offset)), open(p + ".expected.json", "w"), indent=1)
    np.savez_compressed(p + ".bf16.npz", **outs)


def make_gguf_legacy_and_k_quants():
    """One tensor of each remaining legacy / K quant type (Q4_0, Q4_1, Q5_0, Q5_1, Q2_K, Q3_K, Q5_K), sizes chosen so that
    tensor starts fall on different 16-byte phases; expected values from gguf.quants.dequantize."""
    import gguf
    from gguf import GGMLQuantizationType as Q
    rng = np.random.Generator(np.random.Philox(key=79))

    def blocks(nblk, bsz, doffs):
        b = rng.integers(0, 256, size=(nblk, bsz), dtype=np.uint8)
        for off in doffs:
            e = rng.integers(5, 12, size=nblk, dtype=np.uint16)
            m = rng.integers(0, 1024, size=nblk, dtype=np.uint16)
            sgn = rng.integers(0, 2, size=nblk, dtype=np.uint16) << 15
            b[:, off:off + 2] = ((e << 10) | m | sgn).astype("<u2").view(np.uint8).reshape(nblk, 2)
        return b

    p = os.path.join(HERE, "quants_f4.gguf")
    w = gguf.GGUFWriter(p, "llama")
    w.add_tensor("blk.0.attn_q.weight", blocks(3 * 3, 18, [0]).reshape(3, 3 * 18), raw_dtype=Q.Q4_0)            # [3, 96]
    w.add_tensor("blk.0.attn_k.weight", blocks(2 * 5, 20, [0, 2]).reshape(2, 5 * 20), raw_dtype=Q.Q4_1)         # [2, 160]
    w.add_tensor("blk.0.attn_v.weight", blocks(5 * 1, 22, [0]).reshape(5, 22), raw_dtype=Q.Q5_0)                # [5, 32]
    w.add_tensor("blk.0.attn_output.weight", blocks(3 * 2, 24, [0, 2]).reshape(3, 2 * 24), raw_dtype=Q.Q5_1)    # [3, 64]
    w.add_tensor("blk.0.ffn_gate.weight", blocks(3 * 1, 84, [80, 82]).reshape(3, 84), raw_dtype=Q.Q2_K)         # [3, 256]
    w.add_tensor("blk.0.ffn_up.weight", blocks(2 * 2, 110, [108]).reshape(2, 2 * 110), raw_dtype=Q.Q3_K)        # [2, 512]
    w.add_tensor("blk.0.ffn_down.weight", blocks(3 * 1, 176, [0, 2]).reshape(3, 176), raw_dtype=Q.Q5_K)         # [3, 256]
    w.add_tensor("blk.0.attn_norm.weight", rng.standard_normal(12).astype(np.float32))
    w.write_header_to_file(); w.write_kv_data_to_file(); w.write_tensors_to_file(); w.close()
    r = gguf.GGUFReader(p)
    exp, outs = [], {}
    for t in r.tensors:
        shape = [int(x) for x in reversed(t.shape.tolist())]
        exp.append(dict(name=t.name, dtype=t.tensor_type.name, shape=shape, file_offset=int(t.data_offset), nbytes=int(t.n_bytes)))
        f32 = gguf.quants.dequantize(np.array(t.data), t.tensor_type)
        outs[t.name] = bits16(torch.from_numpy(np.ascontiguousarray(f32, dtype=np.float32)).to(torch.bfloat16)).reshape(-1)
    json.dump(dict(tensors=exp, alignment=int(r.alignment), data_offset=int(r.data_offset)), open(p + ".expected.json", "w"), indent=1)
    np.savez_compressed(p + ".bf16.npz", **outs)


def make_gguf_codebook_quants():
    """IQ4_NL, IQ4_XS and MXFP4 (the 4-bit codebook types) written by gguf-py's GGUFWriter; expected values from quants.dequantize."""
    import gguf
    from gguf import GGMLQuantizationType as Q
    rng = np.random.Generator(np.random.Philox(key=80))

    def blocks(nblk, bsz, d_off=None, e8m0_off=None):
        b = rng.integers(0, 256, size=(nblk, bsz), dtype=np.uint8)
        if d_off is not None:
            e = rng.integers(5, 12, size=nblk, dtype=np.uint16)
            m = rng.integers(0, 1024, size=nblk, dtype=np.uint16)
            sgn = rng.integers(0, 2, size=nblk, dtype=np.uint16) << 15
            b[:, d_off:d_off + 2] = ((e << 10) | m | sgn).astype("<u2").view(np.uint8).reshape(nblk, 2)
        if e8m0_off is not None:
            b[:, e8m0_off] = rng.integers(100, 140, size=nblk, dtype=np.uint8)
            b[0, e8m0_off], b[1 % nblk, e8m0_off] = 0, 1  # the two subnormal encodings of the shared exponent
        return b

    p = os.path.join(HERE, "quants_cb.gguf")
    w = gguf.GGUFWriter(p, "llama")
    w.add_tensor("blk.0.attn_q.weight", blocks(3 * 3, 18, d_off=0).reshape(3, 3 * 18), raw_dtype=Q.IQ4_NL)   # [3, 96]
    w.add_tensor("blk.0.ffn_up.weight", blocks(2 * 2, 136, d_off=0).reshape(2, 2 * 136), raw_dtype=Q.IQ4_XS)  # [2, 512]
    w.add_tensor("blk.0.ffn_gate.weight", blocks(5 * 3, 17, e8m0_off=0).reshape(5, 3 * 17), raw_dtype=Q.MXFP4)  # [5, 96]
    w.add_tensor("blk.0.attn_norm.weight", rng.standard_normal(12).astype(np.float32))
    w.write_header_to_file(); w.write_kv_data_to_file(); w.write_tensors_to_file(); w.close()
    r = gguf.GGUFReader(p)
    exp, outs = [], {}
    for t in r.tensors:
        shape = [int(x) for x in reversed(t.shape.tolist())]
        exp.append(dict(name=t.name, dtype=t.tensor_type.name, shape=shape, file_offset=int(t.data_offset), nbytes=int(t.n_bytes)))
        f32 = gguf.quants.dequantize(np.array(t.data), t.tensor_type)
        outs[t.name] = bits16(torch.from_numpy(np.ascontiguousarray(f32, dtype=np.float32)).to(torch.bfloat16)).reshape(-1)
    json.dump(dict(tensors=exp, alignment=int(r.alignment), data_offset=int(r.data_offset)), open(p + ".expected.json", "w"), indent=1)
    np.savez_compressed(p + ".bf16.npz", **outs)


def make_gguf_lattice_quants():
    """The lattice i-quants (IQ2_XXS, IQ2_XS, IQ2_S, IQ3_XXS, IQ3_S, IQ1_S, IQ1_M), the ternary types (TQ1_0, TQ2_0) and NVFP4, written by
    gguf-py's GGUFWriter from random block bytes; expected values from quants.dequantize."""
    import gguf
    from gguf import GGMLQuantizationType as Q
    rng = np.random.Generator(np.random.Philox(key=81))

    def blocks(nblk, bsz, d_off=None):
        b = rng.integers(0, 256, size=(nblk, bsz), dtype=np.uint8)
        if d_off is not None:
            e = rng.integers(5, 12, size=nblk, dtype=np.uint16)
            m = rng.integers(0, 1024, size=nblk, dtype=np.uint16)
            sgn = rng.integers(0, 2, size=nblk, dtype=np.uint16) << 15
            b[:, d_off:d_off + 2] = ((e << 10) | m | sgn).astype("<u2").view(np.uint8).reshape(nblk, 2)
        return b

    def iq1m(nblk):
        b = blocks(nblk, 56)
        d = ((rng.integers(5, 12, size=nblk, dtype=np.uint16) << 10) | rng.integers(0, 1024, size=nblk, dtype=np.uint16)).astype(np.uint16)
        for i in range(4):  # the fp16 super-scale lives in the top nibbles of the four scale words
            b[:, 48 + 2 * i + 1] = (b[:, 48 + 2 * i + 1] & 0x0F) | ((((d >> (4 * i)) & 0xF) << 4).astype(np.uint8))
        return b

    p = os.path.join(HERE, "quants_iq.gguf")
    w = gguf.GGUFWriter(p, "llama")
    w.add_tensor("blk.0.attn_q.weight", blocks(2 * 2, 66, 0).reshape(2, 2 * 66), raw_dtype=Q.IQ2_XXS)      # [2, 512]
    w.add_tensor("blk.0.attn_k.weight", blocks(3 * 1, 74, 0).reshape(3, 74), raw_dtype=Q.IQ2_XS)           # [3, 256]
    w.add_tensor("blk.0.attn_v.weight", blocks(2 * 1, 82, 0).reshape(2, 82), raw_dtype=Q.IQ2_S)            # [2, 256]
    w.add_tensor("blk.0.attn_output.weight", blocks(3 * 1, 98, 0).reshape(3, 98), raw_dtype=Q.IQ3_XXS)     # [3, 256]
    w.add_tensor("blk.0.ffn_gate.weight", blocks(2 * 2, 110, 0).reshape(2, 2 * 110), raw_dtype=Q.IQ3_S)    # [2, 512]
    w.add_tensor("blk.0.ffn_up.weight", blocks(3 * 1, 50, 0).reshape(3, 50), raw_dtype=Q.IQ1_S)            # [3, 256]
    w.add_tensor("blk.0.ffn_down.weight", iq1m(2 * 1).reshape(2, 56), raw_dtype=Q.IQ1_M)                   # [2, 256]
    w.add_tensor("blk.1.attn_q.weight", blocks(3 * 1, 54, 52).reshape(3, 54), raw_dtype=Q.TQ1_0)           # [3, 256]
    w.add_tensor("blk.1.attn_k.weight", blocks(2 * 1, 66, 64).reshape(2, 66), raw_dtype=Q.TQ2_0)           # [2, 256]
    w.add_tensor("blk.1.attn_v.weight", blocks(5 * 3, 36).reshape(5, 3 * 36), raw_dtype=Q.NVFP4)           # [5, 192]
    w.add_tensor("blk.0.attn_norm.weight", rng.standard_normal(12).astype(np.float32))
    w.write_header_to_file(); w.write_kv_data_to_file(); w.write_tensors_to_file(); w.close()
    r = gguf.GGUFReader(p)
    exp, outs = [], {}
    for t in r.tensors:
        shape = [int(x) for x in reversed(t.shape.tolist())]
        exp.append(dict(name=t.name, dtype=t.tensor_type.name, shape=shape, file_offset=int(t.data_offset), nbytes=int(t.n_bytes)))
        f32 = gguf.quants.dequantize(np.array(t.data), t.tensor_type)
        outs[t.name] = bits16(torch.from_numpy(np.ascontiguousarray(f32, dtype=np.float32)).to(torch.bfloat16)).reshape(-1)
    json.dump(dict(tensors=exp, alignment=int(r.alignment), data_offset=int(r.data_offset)), open(p + ".expected.json", "w"), indent=1)
    np.savez_compressed(p + ".bf16.npz", **outs)


def make_sharded():
    from huggingface_hub import save_torch_state_dict
    d = os.path.join(HERE, "sharded")
    shutil.rmtree(d, ignore_errors=True)
    os.makedirs(d)
    g = torch.Generator().manual_seed(5)
    sd = {f"model.layers.{i}.w": (torch.randn(32, 48, generator=g)).to(torch.bfloat16) for i in range(6)}
    sd["lm_head.weight"] = torch.randn(10, 48, generator=g)
    save_torch_state_dict(sd, d, max_shard_size=8000)
    exp = {k: dict(dtype=str(v.dtype), shape=list(v.shape), sha256=sha(raw_bytes(v))) for k, v in sd.items()}
    json.dump(exp, open(os.path.join(d, "expected.json"), "w"), indent=1)


def make_cast_vectors():
    # hand-picked fp32 / fp16 patterns: RNE ties, carries into the exponent, subnormals, +-0, +-inf, max finite
    f32 = np.array([0x00000000, 0x80000000, 0x3F800000, 0x3F808000, 0x3F818000, 0x3F807FFF, 0x3F808001, 0x3FFF8000,
                    0x7F7FFFFF, 0x7F7F7FFF, 0x7F800000, 0xFF800000, 0x00000001, 0x00008000, 0x00018000, 0x007FFFFF,
                    0x80008000, 0x33800000, 0x477FE000, 0xC2F6E979, 0x3EAAAAAB, 0x0000FFFF, 0x7F7F8000], dtype=np.uint32)
    rng = np.random.Generator(np.random.Philox(key=99))
    rnd = rng.integers(0, 1 << 32, size=4096, dtype=np.uint64).astype(np.uint32)
    rnd = rnd[(rnd & 0x7F800000) != 0x7F800000]  # NaN/inf excluded from the torch-pinned set
    f32 = np.concatenate([f32, rnd])
    f32_out = bits16(torch.from_numpy(f32.view(np.float32).copy()).to(torch.bfloat16))
    f16 = np.arange(0, 1 << 16, dtype=np.uint16)
    f16 = f16[((f16 & 0x7C00) != 0x7C00) | ((f16 & 0x03FF) == 0)]  # every fp16 except NaNs
    f16_out = bits16(torch.from_numpy(f16.view(np.float16).copy()).to(torch.bfloat16))
    np.savez_compressed(os.path.join(HERE, "cast_vectors.npz"), f32_in=f32, f32_out=f32_out, f16_in=f16, f16_out=f16_out)


if __name__ == "__main__":
    make_safetensors(); make_gguf(); make_gguf_mixed_quants(); make_gguf_legacy_and_k_quants(); make_gguf_codebook_quants(); make_gguf_lattice_quants(); make_sharded(); make_cast_vectors()
    for r, _, fs in os.walk(HERE):
        for f in sorted(fs):
            p = os.path.join(r, f)
            print(f"{os.path.getsize(p):8d} {os.path.relpath(p, HERE)}")
