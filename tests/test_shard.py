"""Host-side sharding logic on CPU: world_size-2 gloo processes, a stand-in for
the per-rank conversion (the real one needs a B200)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from quickvc_official_b200.shard import HostGather, convert_sharded, shard_range


def test_shard_range_partitions_contiguously():
    for n in (0, 1, 2, 7, 64, 4096):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _fake_infer(unit, mel):
    # deterministic function of the utterance and the target: lets rank 0 verify order and content
    b, _, t = unit.shape
    return (unit.mean(dim=1, keepdim=True).repeat_interleave(320, dim=2) + mel.mean()).reshape(b, 1, 320 * t)


def _worker(rank, world, port, n, t, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        unit = torch.randn(n, 256, t, generator=g)
        mel = torch.randn(1, 80, 40, generator=g)
        out = convert_sharded(_fake_infer, unit, mel)
        if rank == 0:
            ret["ok"] = bool(torch.equal(out, _fake_infer(unit, mel)))
            ret["shape"] = tuple(out.shape)
        else:
            assert out is None
        mine = convert_sharded(_fake_infer, unit, mel, gather=False)
        lo, hi = shard_range(n, world, rank)
        assert torch.equal(mine, _fake_infer(unit[lo:hi], mel))
        # the host gather: every rank copies its own rows into one shared host buffer, two utterances per call
        hg = HostGather(n, 320 * t, name=f"qvc_test_gather_{port}_{n}")
        try:
            got = hg.convert(_fake_infer, unit, mel, chunk=2)
            hg.finish()
            if rank == 0:
                ret["host_ok"] = bool(torch.equal(got, _fake_infer(unit, mel)))
            else:
                assert got is None
        finally:
            hg.close()
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("n", [5, 4, 1])
def test_convert_sharded_two_ranks_gloo(n):
    world, t = 2, 6
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), n, t, ret), nprocs=world, join=True)
        assert ret["ok"] and ret["shape"] == (n, 1, 320 * t)
        assert ret["host_ok"]


def test_convert_sharded_single_process():
    unit = torch.randn(3, 256, 4)
    mel = torch.randn(1, 80, 10)
    assert torch.equal(convert_sharded(_fake_infer, unit, mel), _fake_infer(unit, mel))
