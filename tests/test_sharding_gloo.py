"""CPU, world_size 2 on gloo: the N > 1 host logic of the slice loop (batch sharding, bpp all-reduce,
max-over-ranks timing).  The data path has no collective."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dcae_b200.sharding import gather_symbol_streams, max_over_ranks, reduce_bpp, shard_indices


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_images, pixels_per_image = 5, 512 * 768
        mine = shard_indices(n_images, rank, world)
        # per-image "sum log2 lik" is a deterministic function of the global image id
        per_image = torch.tensor([-(1000.0 + 37.0 * g) for g in range(n_images)], dtype=torch.double)
        local = per_image[mine].sum()
        bpp = reduce_bpp(local.float(), len(mine) * pixels_per_image)
        t = max_over_ranks(10.0 + rank, torch.device("cpu"))
        ret[rank] = (mine, bpp, t)
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_and_reductions():
    world, port = 2, _free_port()
    ret = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert ret[0][0] == [0, 2, 4] and ret[1][0] == [1, 3]
    want = sum(1000.0 + 37.0 * g for g in range(5)) / (5 * 512 * 768)
    assert abs(ret[0][1] - want) < 1e-9 and ret[0][1] == ret[1][1]
    assert ret[0][2] == 11.0 and ret[1][2] == 11.0


def test_shards_cover_everything_once_and_reassemble():
    for world in (1, 2, 4, 8):
        for n in (0, 1, 7, 16):
            shards = [shard_indices(n, r, world) for r in range(world)]
            flat = sorted(i for s in shards for i in s)
            assert flat == list(range(n))
    order = [shard_indices(5, r, 2) for r in range(2)]
    full = torch.arange(5 * 5 * 2).reshape(5, 5, 2, 1, 1).int()        # [slice, image, c, h, w]
    parts = [full[:, o] for o in order]
    assert torch.equal(gather_symbol_streams(parts, order), full)
