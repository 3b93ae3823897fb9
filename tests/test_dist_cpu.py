"""CPU / gloo, world_size 2: the N>1 path — block sharding, logit packing and the single all-gather
— reproduces the concatenation of per-rank results (SURVEY §4 tier 4), including ragged shards."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from llmseg_b200 import dist as lsd


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_local(lo, hi, k_of):
    """Deterministic stand-in for a local forward: image g has K = k_of(g) proposals."""
    sims, ious, best, ks = [], [], [], []
    if hi <= lo:  # empty shard (more ranks than images)
        return torch.zeros((0, 0)), torch.zeros((0, 0)), torch.zeros(0, dtype=torch.int32), []
    kmax = max(k_of(g) for g in range(lo, hi))
    sim = torch.full((hi - lo, kmax), float("-inf"))
    iou = torch.zeros((hi - lo, kmax))
    for i, g in enumerate(range(lo, hi)):
        k = k_of(g)
        gen = torch.Generator().manual_seed(1000 + g)
        sim[i, :k] = torch.rand(k, generator=gen) * 2 - 1
        iou[i, :k] = torch.rand(k, generator=gen)
        ks.append(k)
    best_t = sim.argmax(dim=1).to(torch.int32)
    return sim, iou, best_t, ks


def _worker(rank, world, port, B, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = lsd.init_from_env("gloo")
    assert (r, w) == (rank, world)
    k_of = lambda g: 5 + (g * 7) % 11
    lo, hi = lsd.shard_range(B, rank, world)
    b_max = -(-B // world)
    sim, iou, best, ks = _fake_local(lo, hi, k_of)
    packed = lsd.pack_logits(sim, iou, best, ks, k_max=16, b_max=b_max)
    sims, ious, bests = lsd.unpack_logits(lsd.all_gather_logits(packed), 16)
    # expected: the single-process result over the whole batch
    ok = len(sims) == B
    for g in range(B):
        s1, i1, b1, _ = _fake_local(g, g + 1, k_of)
        k = k_of(g)
        ok &= torch.equal(sims[g], s1[:, :k]) and torch.equal(ious[g], i1[:, :k]) and bests[g] == int(b1[0])
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [5, 1])
def test_allgather_matches_concatenation(B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 32):
        for w in (1, 2, 4, 8):
            spans = [lsd.shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
