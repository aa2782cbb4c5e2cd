"""Multi-rank host logic of the frame-sharded path (ken_burns_effect_b200/utils/shard.py) with two gloo processes on
CPU: pack -> broadcast -> unpack round trip, interleaved pose shards, gather in pose order."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ken_burns_effect_b200.utils import shard


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _fake_common(N=1000, H=12, W=16):
    g = torch.Generator().manual_seed(7)
    return {
        'intHeight': H, 'intWidth': W, 'dblFocal': 512.0, 'dblBaseline': 120,
        'objectDepthrange': (511.25, 3072.5, (3, 4), (9, 2)), 'dblDispmin': 0.0, 'dblDispmax': 120.0,
        'tensorInpaPoints': torch.rand(1, 3, N, generator=g), 'tensorInpaImage': torch.rand(1, 3, N, generator=g),
        'tensorInpaDepth': torch.rand(1, 1, N, generator=g) + 1.0,
    }


def _fake_render(H, W):
    # frame of pose i is filled with (i * 7) % 251 so ordering mistakes are visible
    def fn(poses):
        out = torch.empty(len(poses), H, W, 3, dtype=torch.uint8)
        for j, (idx, _) in enumerate(poses):
            out[j] = (int(idx) * 7) % 251
        return out
    return fn


def _worker(rank, world_size, port, n_poses, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    try:
        common = _fake_common() if rank == 0 else None
        cloud = shard.broadcast_cloud(common, torch.device("cpu"), src=0)
        ref = _fake_common()
        ok = all(torch.equal(cloud[k], ref[k]) for k in ('tensorInpaPoints', 'tensorInpaImage', 'tensorInpaDepth'))
        ok &= all(cloud[k] == ref[k] for k in ('intHeight', 'intWidth', 'dblFocal', 'dblBaseline', 'objectDepthrange'))
        ok &= isinstance(cloud['dblBaseline'], int)
        poses = [(i, 512.0) for i in range(n_poses)]
        frames = shard.render_sharded(poses, _fake_render(cloud['intHeight'], cloud['intWidth']))
        if rank == 0:
            want = np.array([(i * 7) % 251 for i in range(n_poses)], dtype=np.uint8)
            ok &= frames.shape == (n_poses, 12, 16, 3) and bool((frames[:, 0, 0, 0].numpy() == want).all())
            ok &= bool((frames.numpy() == want[:, None, None, None]).all())
        else:
            ok &= frames is None
        idx, local = shard.render_sharded(poses, _fake_render(12, 16), gather=False)
        ok &= idx == list(range(rank, n_poses, world_size)) and local.shape[0] == len(idx)
        # the sync-free exchange: preallocated buffer, header on the host side channel, two effects with different N in a row
        ex = shard.CloudExchange(torch.device("cpu"), capacity_points=1200, src=0)
        for n_pts in (1000, 700, 1500):                       # the last one outgrows the buffer
            c = ex.broadcast(_fake_common(N=n_pts) if rank == 0 else None)
            r2 = _fake_common(N=n_pts)
            ok &= all(torch.equal(c[k], r2[k]) for k in ('tensorInpaPoints', 'tensorInpaImage', 'tensorInpaDepth'))
            ok &= c['objectDepthrange'] == r2['objectDepthrange'] and c['tensorPacked'].shape == (7, n_pts)
        # a cloud that is already packed is sent from where it lies
        if rank == 0:
            packed, hdr = shard.pack_cloud(_fake_common(N=900))
            src_cloud = shard.unpack_cloud(packed, hdr)
        c = ex.broadcast(src_cloud if rank == 0 else None)
        ok &= torch.equal(c['tensorInpaImage'], _fake_common(N=900)['tensorInpaImage'])
        # frames in a shared host segment: every rank fills its block, every rank sees all frames in pose order
        sh = shard.SharedFrames(n_poses, 12, 16, tag=f"kb200test{port}")
        mine = shard.shard_indices(n_poses, rank, world_size)
        blk = sh.block()
        ok &= blk.shape == (len(mine), 12, 16, 3)
        for j, i in enumerate(mine):
            blk[j] = (i * 7) % 251
        dist.barrier()
        ok &= all(bool((sh.frame(i) == (i * 7) % 251).all()) for i in range(n_poses))
        dist.barrier()
        sh.close()
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_poses", [7, 8, 1])
def test_broadcast_shard_gather_world2(n_poses):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_poses, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]


def test_shard_indices_partition():
    for n in (0, 1, 5, 150, 151):
        for R in (1, 2, 3, 8):
            got = sorted(i for r in range(R) for i in shard.shard_indices(n, r, R))
            assert got == list(range(n))
            sizes = [len(shard.shard_indices(n, r, R)) for r in range(R)]
            assert max(sizes) - min(sizes) <= 1


def test_pack_unpack_single_process():
    c = _fake_common()
    packed, hdr = shard.pack_cloud(c)
    assert packed.shape == (7, 1000) and packed.dtype == torch.float32
    u = shard.unpack_cloud(packed, hdr)
    assert torch.equal(u['tensorInpaPoints'], c['tensorInpaPoints']) and u['objectDepthrange'] == c['objectDepthrange']
    assert shard.world() == (0, 1)
