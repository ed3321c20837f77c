"""Single-process, multi-device run of the fused convert + fan-out kernel for `ncu` (never a multi-rank command):
one context over devices 0..N-1 (peer access enabled in kk_open), the checkpoint's parts resident in each GPU's image,
then a few kk_convert_resident rounds.  Usage: python tools/profile_fanout.py <N> [layers]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kukeon_b200 import gpupool  # noqa: E402
from tools import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
layers = int(sys.argv[2]) if len(sys.argv) > 2 else 8
cfg = dict(synth.LLAMA3_8B, layers=layers)
d = f"/dev/shm/kk_prof_llama_{layers}"
if not os.path.exists(os.path.join(d, ".ok")):
    synth.make_llama(d, cfg)
    open(os.path.join(d, ".ok"), "w").write("ok")
with gpupool.Pool(list(range(n))) as pool:
    m = pool.load(d, mode=gpupool.MODE_BROADCAST, fanout=gpupool.FANOUT_P2P, flags=gpupool.LOAD_DEFER)
    m.stage_resident()
    for i in range(6):
        tot, per = m.convert_resident()
        print(f"round {i}: {tot:.3f} ms, launches {[round(x, 3) for x in per]}", flush=True)
    info = m.info()
    sums = {m.checksum(dev, 0, info["pool_bytes"] // 8 * 8) for dev in info["devices"]}
    print("pools identical:", len(sums) == 1, "file GB", info["file_bytes"] / 1e9)
    m.release()
