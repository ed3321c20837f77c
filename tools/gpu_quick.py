"""Torch-free quick hardware check of the paths written after round 1's GPU budget ran out (new quant types, FP8 widening,
8-row transpose tiles, KK_FANOUT_PULL).  Appends one line per check to gpurun_out/quick.log as it goes, so a run that is cut
short still reports what it reached.  The full versions of these checks are tests/test_gpu_quants.py."""
import os
import sys
import tempfile
import time
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "quick.log"), "a")
T0 = time.time()


def say(msg):
    LOG.write(f"[{time.time() - T0:6.2f}s] {msg}\n")
    LOG.flush()
    os.fsync(LOG.fileno())
    print(msg, flush=True)


def main():
    import numpy as np

    from kukeon_b200 import gpupool
    from oracle import oracle
    from tools import synth
    say("imports done")
    G = os.path.join(ROOT, "tests", "golden")
    pool = gpupool.Pool([0], n_staging_buffers=2, staging_buffer_bytes=2 << 20, n_reader_threads=1)
    say("kk_open ok")

    def check(label, path, **kw):
        try:
            shards, recs = oracle.index_path(path)
            oflags = kw.get("flags", 0)
            m = pool.load(path, **kw)
            try:
                exp, plan = oracle.expected_pool(shards, recs, 0, oflags)
                got = m.read(0, 0, len(exp))
                bad = [p["name"] + ":" + recs[i]["dtype"] for i, p in enumerate(plan)
                       if not np.array_equal(got[p["pool_offset"]:p["pool_offset"] + p["nbytes"]], exp[p["pool_offset"]:p["pool_offset"] + p["nbytes"]])]
            finally:
                m.release()
            say(f"{'PASS' if not bad else 'FAIL'} {label}" + (f" differs: {bad[:6]}" if bad else ""))
        except Exception as e:  # noqa: BLE001
            say(f"ERROR {label}: {e!r}")

    check("golden quants_f4 (Q4_0 Q4_1 Q5_0 Q5_1 Q2_K Q3_K Q5_K)", os.path.join(G, "quants_f4.gguf"))
    check("golden quants_cb (IQ4_NL IQ4_XS MXFP4)", os.path.join(G, "quants_cb.gguf"))
    check("golden quants_iq (IQ2_XXS IQ2_XS IQ2_S IQ3_XXS IQ3_S IQ1_S IQ1_M TQ1_0 TQ2_0 NVFP4)", os.path.join(G, "quants_iq.gguf"))
    with tempfile.TemporaryDirectory() as d:
        from tests.test_plan import f4_tensors
        p = os.path.join(d, "f4.gguf")
        synth.write_gguf(p, f4_tensors(hidden=512, ffn=1536, layers=1, vocab=1024), 11)
        check("llama-shaped mix of every new quant type (multi-tile)", p)
        p = os.path.join(d, "fp8.safetensors")
        synth.write_safetensors(p, [("a.weight", "F8_E4M3", [300, 512]), ("b.weight", "F8_E5M2", [129, 65]), ("c.weight", "F8_E4M3", [7])], 23)
        check("FP8 verbatim", p)
        check("FP8 -> bf16 widening", p, flags=gpupool.LOAD_F8_TO_BF16)
        p = os.path.join(d, "gpt2.safetensors")
        synth.make_gpt2(p, n_layer=2, d=96, vocab=301, n_pos=40)
        check("GPT-2 transposes", p, flags=gpupool.LOAD_GPT2_CONV1D_T)
        check("GPT-2 transposes, 4-byte outputs (KEEP_F32)", p, flags=gpupool.LOAD_GPT2_CONV1D_T | gpupool.LOAD_KEEP_F32)
        q = os.path.join(d, "gpt2w.safetensors")
        synth.write_safetensors(q, synth.gpt2_tensors(n_layer=1, d=1032, vocab=50, n_pos=8, dtype="F32"), 3)
        check("transposes, rows wider than one tile (d=1032)", q, flags=gpupool.LOAD_GPT2_CONV1D_T)
        q = os.path.join(d, "gpt2o.safetensors")
        synth.write_safetensors(q, synth.gpt2_tensors(n_layer=1, d=43, vocab=50, n_pos=8, dtype="F16"), 3)
        check("transposes, destination rows off 16 bytes (d=43)", q, flags=gpupool.LOAD_GPT2_CONV1D_T)
        # KK_FANOUT_PULL with 4 virtual ranks on this GPU
        try:
            ld = os.path.join(d, "llama")
            synth.make_llama(ld, dict(hidden=256, ffn=704, layers=2, kv_dim=64, vocab=3000), max_shard_bytes=3_000_000)
            shards, recs = oracle.index_path(ld)
            n = 4
            ms = [pool.load(ld, mode=gpupool.MODE_BROADCAST, fanout=gpupool.FANOUT_PULL, flags=gpupool.LOAD_DEFER, part_index=i, part_count=n) for i in range(n)]
            try:
                ptrs = [m.export_buffer(0, gpupool.BUF_SLICE_PTR) for m in ms]
                for i, m in enumerate(ms):
                    for j in range(n):
                        if j != i:
                            m.peer_attach_buffer(j, gpupool.BUF_SLICE_PTR, ptrs[j])
                for m in ms:
                    m.load_part()
                for m in ms:
                    m.convert_local()
                exp, plan = oracle.expected_pool(shards, recs, 1, 0)
                bad = 0
                for m in ms:
                    got = m.read(0, 0, len(exp))
                    bad += sum(not np.array_equal(got[p["pool_offset"]:p["pool_offset"] + p["nbytes"]], exp[p["pool_offset"]:p["pool_offset"] + p["nbytes"]]) for p in plan)
                say(f"{'PASS' if not bad else 'FAIL'} KK_FANOUT_PULL, 4 virtual ranks ({bad} tensor mismatches)")
            finally:
                for m in ms:
                    m.release()
        except Exception as e:  # noqa: BLE001
            say(f"ERROR KK_FANOUT_PULL: {e!r}")
    try:
        w = max(pool.probe_hbm(0, gpupool.PROBE_WRITE, 4 << 30) for _ in range(3))
        c = max(pool.probe_hbm(0, gpupool.PROBE_COPY, 2 << 30) for _ in range(3))
        say(f"INFO HBM probes: store-only {w:.0f} GB/s, ld/st copy {c:.0f} GB/s (read + write)")
    except Exception as e:  # noqa: BLE001
        say(f"ERROR HBM probe: {e!r}")
    pool.close()
    say("done")


if __name__ == "__main__":
    try:
        main()
    except Exception:  # noqa: BLE001
        say("FATAL " + traceback.format_exc().replace("\n", " | "))
        raise
