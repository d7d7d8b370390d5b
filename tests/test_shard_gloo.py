"""N > 1 host logic on CPU: world_size 2 over gloo (the data path itself needs no collective)."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from damapper_b200 import shard, synth, dazzdb
    from oracle import oracle as orc                     # checker: builds the index both ranks compare
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        contigs, rb, rl = synth.make_config("C1", scale=0.01, seed=11)
        rf = dazzdb.load_block(contigs)
        idx = orc.sort_kmers(orc.HostBlock(*rf), kmer=20, suppress=0)          # (len+2, 16) bytes
        mine = np.ascontiguousarray(idx).view(np.uint8).reshape(-1)
        payload = torch.from_numpy(mine.copy()) if rank == 0 else None
        buf, ln = shard.broadcast_index(dist, torch, payload, "cpu", src=0)
        ok = (ln == mine.size // 16 - 2) and bool((buf.numpy() == mine).all())
        blocks = shard.assign_blocks(7, world, rank)
        got = [None] * world
        dist.all_gather_object(got, blocks)
        flat = sorted(b for bl in got for b in bl)
        ok = ok and flat == list(range(7)) and all(len(set(a) & set(b)) == 0 for i, a in enumerate(got) for b in got[i + 1:])
        q.put((rank, ok, ln))
    finally:
        dist.destroy_process_group()


def test_index_broadcast_and_block_sharding_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res), res
    assert res[0][2] == res[1][2] and res[0][2] > 0


def test_assign_blocks_edges():
    from damapper_b200 import shard
    assert shard.assign_blocks(0, 4, 1) == []
    assert shard.assign_blocks(3, 8, 5) == []
    assert shard.assign_blocks(10, 4, 3) == [3, 7]
    with pytest.raises(ValueError):
        shard.assign_blocks(4, 2, 2)
