"""CPU tier: the loader's page-cache -> staging-slot read paths (kk_loader.cpp FdSet / read_chunk / copy_nt) through tests/emul/kk_read_test —
every KUKEON_GPULOAD_READ mode must deliver the same bytes: long ranges off page and 32-byte boundaries, short ranges, a second pass after
the per-range MADV_DONTNEED, and KK_EIO for a range past the end of the file."""
import json
import os
import shutil
import subprocess
import tempfile

import pytest

_HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(_HERE, "emul", "_build", "kk_read_test")
MODES = ["auto", "pread", "mapped"]


@pytest.fixture(scope="module")
def binary(native):
    subprocess.run(["make", "-C", os.path.join(_HERE, "emul"), "-s"], check=True)
    return BIN


def run(binary, d, policy, mode, *extra):
    r = subprocess.run([binary, d, policy, *extra], env=dict(os.environ, KUKEON_GPULOAD_READ=mode), capture_output=True, text=True, timeout=120)
    doc = json.loads(r.stdout.strip().splitlines()[-1])
    assert r.returncode == 0 and doc["bad"] == 0 and doc["error"] == "", (mode, policy, doc, r.stderr[-500:])
    return doc


def _is_tmpfs(path: str) -> bool:
    return subprocess.run(["stat", "-f", "-c", "%T", path], capture_output=True, text=True).stdout.strip() == "tmpfs"


@pytest.mark.skipif(not os.path.isdir("/dev/shm"), reason="needs a tmpfs")
@pytest.mark.parametrize("mode", MODES)
def test_every_read_mode_delivers_the_file_bytes_from_tmpfs(binary, mode):
    d = tempfile.mkdtemp(prefix="kk_read_", dir="/dev/shm")
    try:
        doc = run(binary, d, "tmpfs", mode)
        assert doc["mapped"] == 1 and doc["mode"] == MODES.index(mode)
        for seed in ("1", "2", "3"):  # other random chunks (1 B .. 2 MiB ranges at arbitrary offsets)
            run(binary, d, "all", mode, seed)
        assert run(binary, d, "none", mode)["mapped"] == 0  # no mapping: every mode falls back to pread
    finally:
        shutil.rmtree(d, ignore_errors=True)


def test_only_tmpfs_shards_are_mapped_by_default(binary, tmp_path):
    doc = run(binary, str(tmp_path), "tmpfs", "auto")
    assert doc["mapped"] == (1 if _is_tmpfs(str(tmp_path)) else 0)
    assert run(binary, str(tmp_path), "all", "mapped")["mapped"] == 1  # a forced mode maps whatever the file system
