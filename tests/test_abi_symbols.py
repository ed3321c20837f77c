"""The C-ABI library loads without a GPU and exports every symbol include/kukeon_gpuload.h declares."""
import ctypes
import os
import re

import pytest

from kukeon_b200 import gpupool

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "kukeon_gpuload.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(kk_[a-z_0-9]+)\s*\(", src)))


def test_header_symbols_all_exported(native):
    syms = declared_symbols()
    assert len(syms) >= 25
    L = ctypes.CDLL(gpupool.lib_path())
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(gpupool.ABI_SYMBOLS) == syms, "gpupool.ABI_SYMBOLS must list exactly the header's functions"


def test_abi_version_and_status_names(native):
    assert native.kk_abi_version() == 1
    assert native.kk_status_name(0) == b"KK_OK"
    assert native.kk_status_name(-6) == b"KK_ECUDA"
    assert native.kk_status_name(-3) == b"KK_EFORMAT"


def test_struct_sizes_match_header():
    # layouts the Go/cgo side would mirror; sizes computed from the header by hand
    assert ctypes.sizeof(gpupool.KKConfig) == 4 + 32 + 4 + 8 + 4 + 4 + 8 + 4 + 4  # with natural padding = 72
    assert ctypes.sizeof(gpupool.KKTensorMeta) == 256 + 4 + 4 + 64 + 4 + 4 + 8 + 8
    assert ctypes.sizeof(gpupool.KKPlacement) == 4 + 4 + 8 + 8 + 4 + 4 + 64 + 8
    assert ctypes.sizeof(gpupool.KKLoadOpts) == 32


def test_null_arguments_are_rejected_not_crashed(native):
    assert native.kk_open(None, None) == -1
    assert b"NULL" in native.kk_last_error()
    assert native.kk_close(None) == -1
    assert native.kk_release(None) == -1
    assert native.kk_index(None, None, None, None) == -1
    assert native.kk_probe_hbm(None, 0, 0, 1 << 20, None) == -1


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(gpupool, "_lib", None)
    monkeypatch.setattr(gpupool, "lib_path", lambda: str(tmp_path / "nope.so"))
    with pytest.raises(ImportError, match="no CPU fallback"):
        gpupool.lib()


@pytest.mark.skipif(__import__("torch").cuda.is_available(), reason="CPU-only behaviour")
def test_no_gpu_means_error_not_fallback(native):
    with pytest.raises(gpupool.ErrCUDA):
        gpupool.Pool([0])


def test_cpu_entry_points_are_thread_safe_and_errors_are_per_thread(native, tmp_path):
    """kukeond serves one goroutine per connection (internal/daemon/server.go:236), so calls arrive on arbitrary OS threads
    at once: 16 threads index and plan good and malformed checkpoints concurrently; every thread must see its own
    kk_last_error (thread-local) and the same records as a single-threaded run."""
    import threading

    from tests import helpers
    from tools import synth
    good = str(tmp_path / "g.safetensors")
    helpers.mixed_safetensors(good)
    gg = str(tmp_path / "g.gguf")
    synth.write_gguf(gg, [("blk.0.attn_q.weight", "Q4_K", [8, 256]), ("blk.0.ffn_up.weight", "Q5_K", [4, 512]), ("blk.0.attn_norm.weight", "F32", [16])], 3)
    bad = str(tmp_path / "bad.safetensors")
    open(bad, "wb").write(b"\xff" * 64)
    want = {good: gpupool.index(good), gg: gpupool.index(gg)}
    want_plan = {p: gpupool.plan_describe(p, mode=gpupool.MODE_BROADCAST, n_parts=3) for p in want}
    errs = []

    def worker(k):
        try:
            for it in range(40):
                p = (good, gg)[(k + it) % 2]
                assert gpupool.index(p) == want[p]
                if it % 4 == 0:
                    assert gpupool.plan_describe(p, mode=gpupool.MODE_BROADCAST, n_parts=3) == want_plan[p]
                if k % 2:  # odd threads interleave failing calls; their message must be theirs
                    missing = str(tmp_path / f"missing-{k}.safetensors")
                    with pytest.raises(gpupool.ErrNotFound, match=f"missing-{k}"):
                        gpupool.index(missing)
                    with pytest.raises(gpupool.ErrFormat):
                        gpupool.index(bad)
        except BaseException as e:  # noqa: BLE001
            errs.append((k, repr(e)))

    th = [threading.Thread(target=worker, args=(k,)) for k in range(16)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs


def test_every_environment_knob_of_the_library_is_documented():
    """The library's getenv() names are measurement / test aids outside the C ABI; DESIGN.md §4 lists them all."""
    import glob
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    names = set()
    for f in glob.glob(os.path.join(root, "kukeon_b200", "csrc", "*.c*")):
        names |= set(re.findall(r'getenv\("([A-Z0-9_]+)"\)', open(f).read()))
    assert names, "no getenv found: the scan is broken"
    design = open(os.path.join(root, "DESIGN.md")).read()
    missing = sorted(n for n in names if n not in design)
    assert not missing, f"undocumented environment knobs: {missing}"
