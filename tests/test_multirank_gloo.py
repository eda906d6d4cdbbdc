"""N>1 host-side logic on CPU: world_size-2 gloo run of the scene sharding and the gradient all-reduce
(the data path itself has no collective)."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stratified_transformer_b200 import parallel


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sizes = [5, 3, 4, 6, 2]
        offset = torch.tensor(sizes).cumsum(0).int()
        xyz = torch.arange(sum(sizes) * 3, dtype=torch.float32).view(-1, 3)
        scene_of_point = torch.repeat_interleave(torch.arange(len(sizes)), torch.tensor(sizes))
        x, off, sid, ids = parallel.shard_batch(xyz, offset, rank, world, scene_of_point)
        assert ids == list(range(rank, len(sizes), world))
        assert off.tolist() == torch.tensor([sizes[i] for i in ids]).cumsum(0).tolist()
        assert sorted(set(sid.tolist())) == ids and x.shape[0] == off[-1]
        # every scene is owned by exactly one rank
        owned = torch.zeros(len(sizes))
        owned[ids] = 1
        dist.all_reduce(owned)
        assert owned.tolist() == [1.0] * len(sizes)
        # gradient all-reduce: one flat collective, in place, mean over ranks
        g1, g2 = torch.full((3, 2), float(rank + 1)), torch.arange(4.0) * (rank + 1)
        parallel.allreduce_gradients([g1, None, g2])
        assert torch.allclose(g1, torch.full((3, 2), 1.5)) and torch.allclose(g2, torch.arange(4.0) * 1.5)
        # the same started early and finished later (one layer's exchange under the next layer's backward)
        g3 = torch.full((5,), float(rank))
        fin = parallel.allreduce_gradients([g3], async_op=True)
        assert callable(fin)
        fin()
        assert torch.allclose(g3, torch.full((5,), 0.5))
        # whole-job throughput = total points / slowest rank
        thr = parallel.global_throughput(100 * (rank + 1), 1.0 + rank)
        assert abs(thr - 300 / 2.0) < 1e-9
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_allreduce():
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert dict(out) == {0: 1, 1: 1}


def test_shard_scenes_single_rank_is_identity():
    assert parallel.shard_scenes(5, 0, 1) == [0, 1, 2, 3, 4]
