import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from kukeon_b200 import gpupool
from oracle import oracle
from tools import synth
import tempfile
d = tempfile.mkdtemp()
f = os.path.join(d, "gpt2.safetensors")
synth.make_gpt2(f, n_layer=1, d=96, vocab=301, n_pos=40)
shards, recs = oracle.index_path(f)
with gpupool.Pool([0], n_staging_buffers=2, staging_buffer_bytes=1 << 20, n_reader_threads=1) as pl:
    m = pl.load(f, flags=1)
    exp, plan = oracle.expected_pool(shards, recs, 0, 1)
    got = m.read(0, 0, len(exp))
    for p in plan:
        a, b = p["pool_offset"], p["pool_offset"] + p["nbytes"]
        if not np.array_equal(got[a:b], exp[a:b]):
            g = got[a:b].view(np.uint16).reshape(p["shape"]); e = exp[a:b].view(np.uint16).reshape(p["shape"])
            bad = np.argwhere(g != e)
            print(p["name"], p["shape"], "transposed" if p["transposed"] else "", "bad", len(bad), "of", g.size)
            rows = sorted(set(bad[:, 0].tolist())); cols = sorted(set(bad[:, 1].tolist())) if bad.shape[1] > 1 else []
            print("  bad dst rows:", rows[:40], "...", len(rows)); print("  bad dst cols:", cols[:40], "...", len(cols))
            r, c = bad[0]; print("  first bad", (r, c), hex(g[r, c]), "want", hex(e[r, c]), "zero?" , int(g[r,c])==0)
            where = np.argwhere(e == g[r, c]); print("  got value appears in expected at", where[:5].tolist())
    print("done", os.environ.get("KUKEON_GPULOAD_LIB", "default"))
    m.release()
