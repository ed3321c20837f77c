"""GPU error-path tests (argument, state and budget errors through the C ABI).  Kept in their own file that sorts after every value-parity
file, so that under `pytest -x` a surprise in an error path can never hide a parity test (round 1's GPUTEST lost six parity cases that way)."""
import pytest

from kukeon_b200 import gpupool
from oracle import oracle
from tests import helpers
from tests.test_gpu_load import assert_pool_matches
from tools import synth

pytestmark = pytest.mark.gpu
MB = 1 << 20


def test_pool_budget_and_error_paths(native, tmp_path):
    p = str(tmp_path / "m.safetensors")
    helpers.mixed_safetensors(p)
    with gpupool.Pool([0], pool_bytes_per_device=4096, n_staging_buffers=2, staging_buffer_bytes=1 * MB, n_reader_threads=1) as pl:
        with pytest.raises(gpupool.ErrNoMemory, match="budget"):
            pl.load(p)
        with pytest.raises(gpupool.ErrNotFound):
            pl.load(str(tmp_path / "missing"))
        with pytest.raises(gpupool.ErrInvalid):
            pl.load(p, mode=9)
        with pytest.raises(gpupool.ErrUnsupported):
            pl.load(p, fanout=gpupool.FANOUT_NVLS)
    with pytest.raises(gpupool.ErrInvalid):
        gpupool.Pool([99])
    with pytest.raises(gpupool.ErrInvalid):
        gpupool.Pool([0, 0])


def test_error_paths_through_the_abi(pool, tmp_path):
    import ctypes as C
    p = str(tmp_path / "m.safetensors")
    helpers.mixed_safetensors(p)
    m = pool.load(p, flags=gpupool.LOAD_KEEP_F32)
    try:
        L = gpupool.lib()
        assert m.placements("b.f32")[0].dtype == "F32" and m.placements("c.f16")[0].dtype == "BF16"
        shards, recs = oracle.index_path(p)
        assert_pool_matches(m, 0, shards, recs, flags=gpupool.LOAD_KEEP_F32)
        with pytest.raises(gpupool.ErrNotFound):
            m.placements("no.such.tensor")
        with pytest.raises(gpupool.ErrInvalid):
            m.read(0, m.info()["pool_bytes"], 16)
        with pytest.raises(gpupool.ErrInvalid):
            m.checksum(0, 4, 16)  # offset must be a multiple of 8
        with pytest.raises(gpupool.ErrInvalid):
            m.export(3)           # device without a pool of this model
        small = C.create_string_buffer(8)
        assert L.kk_export(m._h, 0, None, small, 8) == -9 and b"manifest needs" in L.kk_last_error()   # KK_ERANGE
        assert L.kk_stats(m._h, small, 8) == -9
        with pytest.raises(gpupool.ErrState):
            m.peer_attach(1, b"\0" * 64)  # not a multi-process model
        with pytest.raises(gpupool.ErrState):
            m.convert_local()             # not a RAW model
        m.acquire()
        m.release()
        assert m.info()["refcount"] == 1
    finally:
        m.release()


def test_pull_argument_and_state_errors(pool, tmp_path):
    d = str(tmp_path / "llama")
    synth.make_llama(d, dict(hidden=128, ffn=352, layers=1, kv_dim=32, vocab=500), max_shard_bytes=3_000_000)
    with pytest.raises(gpupool.ErrInvalid, match="one-process-per-GPU"):
        pool.load(d, mode=gpupool.MODE_BROADCAST, fanout=gpupool.FANOUT_PULL)
    with pytest.raises(gpupool.ErrInvalid):
        pool.load(d, mode=gpupool.MODE_SCATTER, fanout=gpupool.FANOUT_PULL, part_index=0, part_count=2)
    f = str(tmp_path / "gpt2.safetensors")
    synth.make_gpt2(f, n_layer=1, d=96, vocab=301, n_pos=40)
    with pytest.raises(gpupool.ErrUnsupported, match="transposing"):
        pool.load(f, mode=gpupool.MODE_BROADCAST, fanout=gpupool.FANOUT_PULL, flags=gpupool.LOAD_GPT2_CONV1D_T | gpupool.LOAD_DEFER, part_index=0, part_count=2)
    m = pool.load(d, mode=gpupool.MODE_BROADCAST, fanout=gpupool.FANOUT_PULL, flags=gpupool.LOAD_DEFER, part_index=0, part_count=2)
    try:
        m.load_part()
        with pytest.raises(gpupool.ErrState, match="not attached"):
            m.convert_local()  # the other rank's slice was never attached: refuse, do not leave half a pool marked loaded
        assert not m.info()["loaded"]
        with pytest.raises(gpupool.ErrState):
            m.peer_attach_buffer(1, gpupool.BUF_RAW, b"\0" * 64)
        with pytest.raises(gpupool.ErrState, match="no attached buffer"):
            m.probe_peer(1, gpupool.BUF_SLICE)
        with pytest.raises(gpupool.ErrInvalid):
            m.probe_peer(1, 99)
    finally:
        m.release()
