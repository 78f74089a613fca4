"""CPU, world_size 2, gloo: the N>1 host logic (sharding, the single conditioning broadcast, gather order)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from textflux_b200 import dist as td
    try:
        g = torch.Generator().manual_seed(7)
        prompt = torch.randn(1, 16, 128, generator=g).to(torch.bfloat16)
        pooled = torch.randn(1, 32, generator=g).to(torch.bfloat16)
        sig = torch.linspace(1, 0, 5)
        if rank != 0:  # other ranks start from garbage
            prompt, pooled, sig = torch.zeros_like(prompt), torch.ones_like(pooled), torch.zeros_like(sig)
        n0 = getattr(dist.broadcast, "__wrapped__", None)
        p2, c2, s2, gs = td.broadcast_conditioning(prompt, pooled, sig, 30.0 if rank == 0 else -1.0)
        g = torch.Generator().manual_seed(7)
        assert torch.equal(p2, torch.randn(1, 16, 128, generator=g).to(torch.bfloat16))
        assert torch.equal(c2, torch.randn(1, 32, generator=g).to(torch.bfloat16))
        assert torch.equal(s2, torch.linspace(1, 0, 5)) and gs == 30.0
        B = 5
        mine = td.shard_indices(B, rank, world)
        lat = torch.stack([torch.full((4, 8), float(b)) for b in mine]).to(torch.bfloat16)
        allg = td.gather_latents(lat, B)
        if rank == 0:
            assert allg.shape == (B, 4, 8)
            assert [int(allg[b, 0, 0]) for b in range(B)] == list(range(B))
        else:
            assert allg is None
        assert td.max_over_ranks(float(rank + 1), "cpu") == float(world)
        q.put((rank, mine, "ok"))
    except Exception as e:  # noqa
        q.put((rank, None, repr(e)))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_single_broadcast():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    res.sort()
    assert all(r[2] == "ok" for r in res), res
    assert res[0][1] == [0, 2, 4] and res[1][1] == [1, 3]


def test_shard_indices_cover_batch_once():
    from textflux_b200.dist import shard_indices
    for B in (1, 4, 8, 13):
        for W in (1, 2, 4, 8):
            seen = sorted(i for r in range(W) for i in shard_indices(B, r, W))
            assert seen == list(range(B))
