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


class _FakeModel:
    """CPU stand-in with the forward's output contract: image g (global id carried in images[:, 0, 0, 0]) has
    K = 5 + (g*7)%11 proposals and similarity rows seeded by (g, conversation)."""
    device = torch.device("cpu")

    def model_forward(self, images, images_clip, input_ids, offset, sam_segs_list, **kw):
        B = images.shape[0]
        off = [int(v) for v in offset.tolist()]
        assert len(off) == B + 1 and off[0] == 0 and off[-1] == input_ids.shape[0] and len(sam_segs_list) == B
        ks = [int(s.shape[0]) for s in sam_segs_list]
        kmax = max(ks)
        sim = torch.full((off[-1], kmax), float("-inf"))
        iou = torch.zeros((B, kmax))
        for i in range(B):
            g = int(images[i, 0, 0, 0])
            assert int(images_clip[i, 0, 0, 0]) == g
            for c in range(off[i], off[i + 1]):
                assert int(input_ids[c, 0]) == g          # the conversation travelled with its image
                gen = torch.Generator().manual_seed(1000 + 10 * g + (c - off[i]))
                sim[c, :ks[i]] = torch.rand(ks[i], generator=gen) * 2 - 1
            iou[i, :ks[i]] = torch.rand(ks[i], generator=torch.Generator().manual_seed(5000 + g))
        first = sim[off[:-1]]
        return {"pred_similarity": [sim[off[i]:off[i + 1], :ks[i]] for i in range(B)],
                "pred_iou": [iou[i:i + 1, :ks[i]] for i in range(B)], "similarity_padded": first,
                "similarity_all": sim, "iou_padded": iou, "best_index": first.argmax(dim=1).to(torch.int32)}


def _global_inputs(B, convs):
    off = [0]
    for c in convs:
        off.append(off[-1] + c)
    ids = torch.cat([torch.full((c, 4), g, dtype=torch.int64) for g, c in enumerate(convs)]) if B else torch.zeros((0, 4), dtype=torch.int64)
    img = torch.arange(B, dtype=torch.float32).view(B, 1, 1, 1).expand(B, 3, 2, 2).contiguous()
    return {"images": img, "images_clip": img.clone(), "input_ids": ids, "labels": ids.clone(),
            "attention_masks": torch.ones_like(ids, dtype=torch.bool), "offset": torch.tensor(off),
            "sam_segs_list": [torch.zeros(5 + (g * 7) % 11, 2, 2) for g in range(B)], "inference": True}


def _dp_worker(rank, world, port, convs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    lsd.init_from_env("gloo")
    B = len(convs)
    inp = _global_inputs(B, convs)
    out = lsd.DataParallelLisa(_FakeModel(), k_max=16).forward(inp)
    ref = _FakeModel().model_forward(**inp)          # the single-process result over the whole batch
    ok = len(out["pred_similarity"]) == B
    for g in range(B):
        ok &= torch.equal(out["pred_similarity"][g], ref["pred_similarity"][g])
        ok &= torch.equal(out["pred_iou"][g], ref["pred_iou"][g])
        ok &= out["best_index"][g] == int(ref["best_index"][g])
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("convs", [[1], [1, 1, 1], [2, 1, 3], [1, 1, 1, 1, 1]])
def test_data_parallel_forward_ragged_and_multi_conversation(convs):
    """DataParallelLisa end to end over gloo, world 2: a global batch SMALLER than the world (one rank owns no image
    and must still join the all-gather), odd batches, and images with several conversations sharded through
    `offset` — all equal to the single-process result."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, convs, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(0, True), (1, True)]


def test_async_all_gather_single_process():
    t = torch.arange(6.0).view(2, 3)
    out, work = lsd.all_gather_logits(t, async_op=True)
    assert work is None and torch.equal(out, t)
