"""Torch-free timing of the transposing load on GPT-2-small (resident image, CUDA-event timed inside the library)."""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from kukeon_b200 import gpupool  # noqa: E402
from tools import synth  # noqa: E402

out = {}
with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as d:
    p = os.path.join(d, "gpt2.safetensors")
    t0 = time.time()
    nbytes = synth.make_gpt2(p)
    out["synth_s"] = time.time() - t0
    pool = gpupool.Pool([0])
    for name, flags in (("transposed_bf16", gpupool.LOAD_GPT2_CONV1D_T), ("plain_cast_bf16", 0), ("transposed_keep_f32", gpupool.LOAD_GPT2_CONV1D_T | gpupool.LOAD_KEEP_F32)):
        m = pool.load(p, flags=flags | gpupool.LOAD_DEFER)
        try:
            m.stage_resident()
            for _ in range(3):
                m.convert_resident()
            runs = [m.convert_resident() for _ in range(20)]
            ms = sorted(t for t, _ in runs)
            st = m.stats()
            alg = st["parts"][0]["src_bytes"] + st["parts"][0]["out_bytes"]
            out[name] = {"ms_median": ms[len(ms) // 2], "ms_min": ms[0], "algorithmic_bytes": alg, "GBps_at_median": alg / (ms[len(ms) // 2] / 1e3) / 1e9,
                         "launches": len(runs[0][1])}
        finally:
            m.release()
    pool.close()
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = json.load(open(pk)).get("hbm_gbs")
    for v in out.values():
        if isinstance(v, dict):
            v["frac_of_copy_peak"] = v["GBps_at_median"] / peak
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", os.environ.get("KK_QUICK_OUT", "gpt2_quick.json")), "w"), indent=1)
print(json.dumps(out))
