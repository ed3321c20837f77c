"""Host-side mirror of the Go API the north_star adds to kukeon's `internal/modelhub`:
`Pull`, `Load`, `Mount` plus the per-Cell acquire/release hooks.

None of these exist in the reference; the shapes below follow the seams they would
sit next to so that the Go shim in INTEGRATION.md is a line-for-line transliteration:

* `Pull`   — local-path resolution + tensor index (nearest analogue: ctr image pull, internal/ctr/image.go:91).
* `Load`   — one call into the C ABI per checkpoint; N concurrent callers share one load
             (runner.StartCell call site, internal/controller/runner/start.go:785-790).
* `Mount`  — stages `<cell metadata dir>/gpupool/{manifest.json,ipc.handle}` atomically
             (internal/metadata/metadata.go:105-140) and returns the OCI bind mount
             (`bindVolumeMount`, internal/ctr/spec.go:526-543) and `KUKEON_GPUPOOL_*` env entries
             (`kukeonDefaultEnv`, internal/ctr/spec.go:464-482) a `ctr.BuildOption` would add.

All data movement happens in libkukeon_gpuload.so; this file never touches tensor bytes.
"""
from __future__ import annotations

import json
import os
import tempfile
from dataclasses import dataclass, field
from typing import Dict, List, Optional

from . import gpupool
from .gpupool import FANOUT_NONE, FANOUT_P2P, MODE_BROADCAST, MODE_SCATTER, MODE_SINGLE  # noqa: F401

CONTAINER_GPUPOOL_DIR = "/run/kukeon/gpupool"  # bind-mount target inside the agent container
ENV_MANIFEST = "KUKEON_GPUPOOL_MANIFEST"
ENV_IPC_HANDLE = "KUKEON_GPUPOOL_IPC_HANDLE"
ENV_DEVICE = "KUKEON_GPUPOOL_DEVICE"


@dataclass
class ModelRef:
    """Result of Pull: where the checkpoint lives and what is in it."""
    path: str
    shards: List[str]
    tensors: List[dict]

    @property
    def file_bytes(self) -> int:
        return sum(t["nbytes"] for t in self.tensors)


def Pull(path: str) -> ModelRef:
    """Resolve a local checkpoint (directory with model.safetensors.index.json / model.safetensors /
    *.gguf, or a single file) and index it.  No network: "pull" is local-path only."""
    return ModelRef(path=os.path.realpath(path), shards=gpupool.index_shards(path), tensors=gpupool.index(path))


def Load(pool: gpupool.Pool, ref: ModelRef | str, mode: int = MODE_SINGLE, fanout: int = FANOUT_P2P, flags: int = 0,
         part_index: int = 0, part_count: int = 0) -> gpupool.Model:
    """Make the checkpoint resident in the pool(s).  Returns a refcounted Model; a second Load of the same
    checkpoint returns the same resident copy with its count bumped."""
    path = ref.path if isinstance(ref, ModelRef) else ref
    return pool.load(path, mode=mode, fanout=fanout, flags=flags, part_index=part_index, part_count=part_count)


@dataclass
class MountSpec:
    """What a `ctr.WithGPUWeights(...)` BuildOption would append to the container's OCI spec."""
    mounts: List[dict] = field(default_factory=list)
    env: List[str] = field(default_factory=list)
    host_dir: str = ""
    devices: List[dict] = field(default_factory=list)        # OCI linux.devices
    device_cgroup: List[dict] = field(default_factory=list)  # OCI linux.resources.devices allow rules


def device_nodes(devices, stat=os.stat) -> tuple:
    """"next" row f2: the NVIDIA character devices an agent container needs to open a CUDA-IPC
    handle — /dev/nvidiactl, /dev/nvidia-uvm, /dev/nvidia-uvm-tools and /dev/nvidia<N> for each exported device — as
    OCI `linux.devices` entries plus the matching device-cgroup allow rules (the reference's spec builder emits
    neither today: internal/ctr/spec.go:218-380).  Nodes that do not exist on the host are skipped."""
    import stat as st_mod
    paths = ["/dev/nvidiactl", "/dev/nvidia-uvm", "/dev/nvidia-uvm-tools"] + [f"/dev/nvidia{d}" for d in sorted(set(devices))]
    devs, rules = [], []
    for p in paths:
        try:
            s = stat(p)
        except OSError:
            continue
        if not st_mod.S_ISCHR(s.st_mode):
            continue
        major, minor = os.major(s.st_rdev), os.minor(s.st_rdev)
        devs.append({"path": p, "type": "c", "major": major, "minor": minor, "fileMode": 0o666, "uid": 0, "gid": 0})
        rules.append({"allow": True, "type": "c", "major": major, "minor": minor, "access": "rw"})
    return devs, rules


def _atomic_write(path: str, data: bytes, mode: int = 0o644) -> None:
    d = os.path.dirname(path)
    fd, tmp = tempfile.mkstemp(prefix=".meta-", suffix=".tmp", dir=d)
    try:
        os.fchmod(fd, mode)
        os.write(fd, data)
        os.fsync(fd)
    finally:
        os.close(fd)
    os.rename(tmp, path)
    try:
        dfd = os.open(d, os.O_RDONLY)
        os.fsync(dfd)
        os.close(dfd)
    except OSError:
        pass


def _env_suffix(name: str) -> str:
    """`llama-3.8b` -> `_LLAMA_3_8B` (POSIX environment names: upper-case letters, digits, underscore)."""
    return "_" + "".join(c.upper() if c.isalnum() else "_" for c in name)


def Mount(model: gpupool.Model, device: int, container_dir: str, with_devices: bool = False, stat=os.stat, name: str = "",
          target: str = CONTAINER_GPUPOOL_DIR) -> MountSpec:
    """Export `device`'s pool for one agent container: write the manifest + IPC handle under
    `<container_dir>/gpupool/` and describe the read-only bind mount and env that expose them.

    `name` (the `models[].name` of the manifest schema, kukeon_b200/schema.py) gives the model its own sub-directory and
    env names — `<container_dir>/gpupool/<name>/`, `<target>/<name>/`, `KUKEON_GPUPOOL_MANIFEST_<NAME>` — so that one container can
    attach several models; without it the single-model layout above is used."""
    if name and (name in (".", "..") or "/" in name or "\0" in name):
        raise ValueError(f"model name {name!r} cannot be used as a directory name")
    handle, manifest = model.export(device)
    host_dir = os.path.join(container_dir, "gpupool", name) if name else os.path.join(container_dir, "gpupool")
    dest = f"{target.rstrip('/')}/{name}" if name else target
    sfx = _env_suffix(name) if name else ""
    os.makedirs(host_dir, mode=0o750, exist_ok=True)
    _atomic_write(os.path.join(host_dir, "manifest.json"), json.dumps(manifest, separators=(",", ":")).encode())
    _atomic_write(os.path.join(host_dir, "ipc.handle"), handle, 0o640)
    devs, rules = device_nodes([device], stat) if with_devices else ([], [])
    return MountSpec(
        mounts=[{"destination": dest, "type": "bind", "source": host_dir, "options": ["rbind", "ro"]}],
        env=[f"{ENV_MANIFEST}{sfx}={dest}/manifest.json", f"{ENV_IPC_HANDLE}{sfx}={dest}/ipc.handle", f"{ENV_DEVICE}{sfx}={device}"],
        host_dir=host_dir, devices=devs, device_cgroup=rules,
    )


def merge_mounts(specs: List[MountSpec]) -> MountSpec:
    """What the container's BuildOptions add up to when several models are mounted: all mounts and env entries, device nodes and
    cgroup rules de-duplicated (two models on the same GPU need /dev/nvidia0 once)."""
    out = MountSpec()
    for s in specs:
        out.mounts += s.mounts
        out.env += s.env
        for d in s.devices:
            if d not in out.devices:
                out.devices.append(d)
        for r in s.device_cgroup:
            if r not in out.device_cgroup:
                out.device_cgroup.append(r)
    dests = [m["destination"] for m in out.mounts]
    if len(set(dests)) != len(dests):
        raise ValueError(f"two models would be mounted at the same container path: {sorted(d for d in dests if dests.count(d) > 1)[0]}")
    return out


class CellHooks:
    """Per-Cell reference counting ("N concurrent agent Sessions" == N Cells).
    acquire() in StartCell; release() from KillCell / StopCell / DeleteCell — idempotent per cell because
    those teardown paths overlap in the reference (markCellFailed calls KillCell, runner/start.go:192-242)."""

    def __init__(self, model: gpupool.Model):
        self.model = model
        self._cells: Dict[str, bool] = {}

    def start_cell(self, cell_id: str) -> None:
        if self._cells.get(cell_id):
            return
        self.model.acquire()
        self._cells[cell_id] = True

    def stop_cell(self, cell_id: str) -> None:
        if self._cells.pop(cell_id, None):
            self.model.release()

    @property
    def active(self) -> int:
        return len(self._cells)
