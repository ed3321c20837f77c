"""Host-side "next" rows: pool registry / restart invalidation (f1) and OCI device exposure (f2).  CPU only."""
import json
import os
import stat as st_mod
import types

import pytest

from kukeon_b200 import modelhub, registry


def make_shards(tmp_path, n=2):
    out = []
    for i in range(n):
        p = tmp_path / f"model-{i}.safetensors"
        p.write_bytes(b"x" * (10 + i))
        out.append(str(p))
    return out


def test_registry_roundtrip_and_same_epoch_keeps_entries(tmp_path):
    run = str(tmp_path / "run")
    shards = make_shards(tmp_path)
    r = registry.PoolRegistry(run, epoch="boot:1:100")
    assert r.reconcile() == []
    r.record("k1", "/models/a", 1, 123, 456, [1, 0], shards, mount_dir=str(tmp_path / "cellA" / "gpupool"))
    r.record("k1", "/models/a", 1, 123, 456, [1, 0], shards, mount_dir=str(tmp_path / "cellB" / "gpupool"))
    doc = json.load(open(os.path.join(run, registry.FILE_NAME)))
    assert doc["apiVersion"] == registry.API_VERSION and doc["epoch"] == "boot:1:100"
    assert doc["models"][0]["devices"] == [0, 1] and len(doc["models"][0]["mounts"]) == 2
    assert st_mod.S_IMODE(os.stat(os.path.join(run, registry.FILE_NAME)).st_mode) == 0o640
    # same process re-reading its own file (config reload): nothing is stale
    r2 = registry.PoolRegistry(run, epoch="boot:1:100")
    assert r2.reconcile() == [] and list(r2.entries) == ["k1"]
    r2.drop_mount("k1", str(tmp_path / "cellA" / "gpupool"))
    assert len(r2.entries["k1"].mounts) == 1
    r2.forget("k1")
    assert json.load(open(os.path.join(run, registry.FILE_NAME)))["models"] == []


def test_restarted_daemon_invalidates_pools_of_the_previous_instance(tmp_path):
    run = str(tmp_path / "run")
    shards = make_shards(tmp_path)
    staged = tmp_path / "cell" / "agent" / "gpupool"
    os.makedirs(staged)
    (staged / "ipc.handle").write_bytes(b"\0" * 64)
    (staged / "manifest.json").write_text("{}")
    other = tmp_path / "cell" / "agent" / "not-ours"
    os.makedirs(other)
    old = registry.PoolRegistry(run, epoch="boot:1:100")
    old.reconcile()
    old.record("k1", "/models/a", 1, 1, 2, [0], shards, mount_dir=str(staged))
    old.record("k2", "/models/b", 0, 3, 4, [1], shards, mount_dir=str(other))  # a path we did not stage: never deleted
    new = registry.PoolRegistry(run, epoch="boot:2:999")  # the daemon came back as another process
    stale = new.reconcile()
    assert sorted(e.key for e in stale) == ["k1", "k2"]
    assert not staged.exists(), "the dead pool's IPC handle must not stay visible to agent containers"
    assert other.exists()
    assert new.entries == {} and json.load(open(os.path.join(run, registry.FILE_NAME)))["epoch"] == "boot:2:999"


def test_restart_also_removes_the_staged_directories_of_named_models(tmp_path):
    """Multi-model layout: Mount stages into <container>/gpupool/<name>.  After a daemon restart those directories — dead ipc.handle and
    manifest.json — must go too (reconcile only matched a basename of exactly 'gpupool'); a look-alike elsewhere stays."""
    run = str(tmp_path / "run")
    shards = make_shards(tmp_path)
    named = tmp_path / "cell" / "agent" / "gpupool" / "llama"
    os.makedirs(named)
    (named / "ipc.handle").write_bytes(b"\0" * 64)
    sibling = tmp_path / "cell" / "agent" / "gpupool" / "other-model"
    os.makedirs(sibling)
    lookalike = tmp_path / "cell" / "agent" / "data" / "llama"
    os.makedirs(lookalike)
    old = registry.PoolRegistry(run, epoch="boot:1:100")
    old.reconcile()
    old.record("k1", "/models/a", 1, 1, 2, [0], shards, mount_dir=str(named))
    old.record("k2", "/models/b", 1, 1, 2, [0], shards, mount_dir=str(lookalike))
    new = registry.PoolRegistry(run, epoch="boot:2:999")
    assert sorted(e.key for e in new.reconcile()) == ["k1", "k2"]
    assert not named.exists() and sibling.exists() and lookalike.exists()


def test_changed_on_disk_and_corrupt_file(tmp_path):
    run = str(tmp_path / "run")
    shards = make_shards(tmp_path)
    r = registry.PoolRegistry(run, epoch="e")
    r.reconcile()
    r.record("k", "/m", 0, 1, 1, [0], shards)
    assert not r.changed_on_disk("k")
    open(shards[0], "ab").write(b"more")
    assert r.changed_on_disk("k")
    os.remove(shards[1])
    assert r.changed_on_disk("k")
    open(os.path.join(run, registry.FILE_NAME), "w").write("{not json")
    with pytest.raises(registry.RegistryError):
        registry.PoolRegistry(run, epoch="e").reconcile()


def test_daemon_epoch_changes_with_the_process():
    a, b = registry.daemon_epoch(), registry.daemon_epoch(1)
    assert a != b and a.count(":") == 2 and a == registry.daemon_epoch()


def test_oci_device_nodes_golden_shape():
    """Same style as the reference's spec tests (internal/ctr/spec_test.go:903-957): build, then assert the exact
    OCI fields.  A fake stat stands in for /dev."""
    table = {"/dev/nvidiactl": (195, 255), "/dev/nvidia-uvm": (511, 0), "/dev/nvidia3": (195, 3), "/dev/nvidia0": (195, 0)}

    def fake_stat(p):
        if p not in table:
            raise FileNotFoundError(p)
        return types.SimpleNamespace(st_mode=st_mod.S_IFCHR | 0o666, st_rdev=os.makedev(*table[p]))

    devs, rules = modelhub.device_nodes([3, 0, 3], stat=fake_stat)
    assert [d["path"] for d in devs] == ["/dev/nvidiactl", "/dev/nvidia-uvm", "/dev/nvidia0", "/dev/nvidia3"]  # uvm-tools absent: skipped
    assert devs[0] == {"path": "/dev/nvidiactl", "type": "c", "major": 195, "minor": 255, "fileMode": 0o666, "uid": 0, "gid": 0}
    assert rules[3] == {"allow": True, "type": "c", "major": 195, "minor": 3, "access": "rw"}
    # a regular file squatting on the name is not passed through as a device
    devs2, _ = modelhub.device_nodes([0], stat=lambda p: types.SimpleNamespace(st_mode=st_mod.S_IFREG | 0o644, st_rdev=0))
    assert devs2 == []
