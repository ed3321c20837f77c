"""N>1 host-side logic on CPU: two ranks (gloo, world_size 2) each plan their own part of a BROADCAST /
SCATTER load through the C ABI, emulate it with the oracle's arithmetic, and exchange results — the parts
must partition the pool exactly (broadcast) or each rank must own its slice (scatter).  The GPU data path for
N>1 (fused P2P stores) is covered by tests/test_gpu_multi.py on a multi-GPU box."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ckpt, mode, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from kukeon_b200 import gpupool
        from oracle import oracle
        from tests import helpers
        shards, recs = oracle.index_path(ckpt)
        plan = gpupool.plan_describe(ckpt, mode=mode, n_parts=world, chunk_bytes=1 << 20)
        layout = plan["layouts"][rank if mode == gpupool.MODE_SCATTER else 0]
        pool, mask = helpers.emulate_part(plan, rank, layout["pool_bytes"])
        if mode == gpupool.MODE_BROADCAST:
            exp, pl = oracle.expected_pool(shards, recs, mode)
            # "all-gather": every rank contributes the bytes it produced; sum of masks must be exactly 1 on tensor bytes
            cover = torch.from_numpy(mask.astype(np.int32))
            dist.all_reduce(cover)
            acc = torch.from_numpy(np.where(mask, pool, 0).astype(np.int32))
            dist.all_reduce(acc)
            want_mask = helpers.expected_mask(pl, len(exp))
            assert (cover.numpy() == want_mask.astype(np.int32)).all(), "parts must partition the pool"
            assert (acc.numpy().astype(np.uint8) == exp).all(), "gathered pool must equal the oracle pool"
            src = torch.tensor([plan["parts"][rank]["src_bytes"]], dtype=torch.int64)
            dist.all_reduce(src)
            assert int(src.item()) == plan["file_bytes"]
        else:
            exp, pl = oracle.expected_pool(shards, recs, mode, 0, world, rank)
            assert (mask == helpers.expected_mask(pl, len(exp))).all()
            assert (pool == exp).all()
            sizes = [None] * world
            dist.all_gather_object(sizes, layout["pool_bytes"])
            assert len(set(sizes)) == 1, "equal slices -> equal pool sizes on every rank"
            # same load with the NVLink row exchange: this rank ingests whole rows of the row-parallel tensors and deals
            # column slices to every rank; summing what all ranks dealt to pool r must give rank r's oracle pool
            xplan = gpupool.plan_describe(ckpt, mode=mode, flags=gpupool.LOAD_SCATTER_EXCHANGE, n_parts=world, chunk_bytes=1 << 20)
            n = layout["pool_bytes"]
            ex = {r: (np.zeros(n, np.uint8), np.zeros(n, bool)) for r in range(world)}
            own, own_mask = helpers.emulate_part(xplan, rank, n, exchange=ex)
            ex[rank][0][own_mask] = own[own_mask]
            ex[rank][1][own_mask] = True
            for r in range(world):
                data = torch.from_numpy(np.where(ex[r][1], ex[r][0], 0).astype(np.int32))
                cover = torch.from_numpy(ex[r][1].astype(np.int32))
                dist.all_reduce(data)
                dist.all_reduce(cover)
                if r == rank:
                    assert (cover.numpy() == helpers.expected_mask(pl, len(exp)).astype(np.int32)).all(), "every pool byte dealt exactly once"
                    assert (data.numpy().astype(np.uint8) == exp).all()
            src = torch.tensor([xplan["parts"][rank]["src_bytes"]], dtype=torch.int64)
            dist.all_reduce(src)
            repl = sum(r_["nbytes"] for r_ in recs if oracle.slice_dim(r_, world) is None)
            assert int(src.item()) == xplan["file_bytes"] + (world - 1) * repl, "row-parallel bytes are read by exactly one rank"
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", [1, 2])
def test_two_rank_plans_partition_the_pool(native, coracle, tmp_path, mode):
    from tools import synth
    d = str(tmp_path / "llama")
    synth.make_llama(d, dict(hidden=128, ffn=352, layers=2, kv_dim=64, vocab=512), max_shard_bytes=600_000)
    out = str(tmp_path / "out")
    os.makedirs(out)
    mp.spawn(_worker, args=(2, _free_port(), d, mode, out), nprocs=2, join=True)
    assert sorted(os.listdir(out)) == ["ok0", "ok1"]
