"""world_size-2 gloo test (CPU) of the multi-GPU host logic: weight broadcast, request
sharding, max-over-ranks timing reduction (mlx_vlm_b200/parallel.py)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mlx_vlm_b200.parallel import broadcast_packed, broadcast_weights, max_over_ranks, shard_requests
    g = torch.Generator().manual_seed(100 + rank)  # different init on every rank
    w = {"b": torch.randn(7, 5, generator=g), "a": torch.randn(33, generator=g).to(torch.bfloat16)}
    nbytes = broadcast_weights(w, src=0)
    ref = torch.Generator().manual_seed(100)
    want_b = torch.randn(7, 5, generator=ref)
    want_a = torch.randn(33, generator=ref).to(torch.bfloat16)
    ok = torch.equal(w["b"], want_b) and torch.equal(w["a"], want_a) and nbytes == 7 * 5 * 4 + 33 * 2
    # the packed arena: every weight is a view of ONE flat buffer -> one broadcast moves them all
    flat_w = torch.randn(1000, generator=g).to(torch.bfloat16)
    views = [flat_w[0:128].view(8, 16), flat_w[256:300]]
    nb = broadcast_packed(flat_w, src=0)
    want_flat = torch.randn(1000, generator=ref).to(torch.bfloat16)
    ok = ok and nb == 2000 and torch.equal(views[0], want_flat[0:128].view(8, 16)) and torch.equal(views[1], want_flat[256:300])
    mine = shard_requests(11, world, rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    flat = sorted(i for part in gathered for i in part)
    ok = ok and flat == list(range(11))
    mx = max_over_ranks([1.0 + rank, 5.0 - rank])
    ok = ok and mx == [float(world), 5.0]
    # C5 host logic: 5 requests over 2 replicas, each rank drives its own BatchGenerator; the tokens
    # every rank sees must equal the requests run alone (deterministic stand-in engine)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from test_batch_host import _alone, _fake_model
    from mlx_vlm_b200.generate_batch import BatchGenerator
    from mlx_vlm_b200.parallel import generate_sharded
    prompts = [[5, 6, 7, 8, 9], [1, 2], [3, 3, 3], [11] * 9, [4, 2]]
    maxes = [7, 12, 1, 9, 3]

    def make():
        model, proc = _fake_model()
        return BatchGenerator(model, proc, completion_batch_size=2, prefill_batch_size=2, decode_slice=4)
    got = generate_sharded(make, prompts, maxes)
    ok = ok and got == [_alone(p, m)[0] for p, m in zip(prompts, maxes)]

    def make_lockstep():   # the same through the lock-step batched decoder's host logic
        model, proc = _fake_model(lockstep=True)
        return BatchGenerator(model, proc, completion_batch_size=3, prefill_batch_size=2, decode_slice=4)
    got = generate_sharded(make_lockstep, prompts, maxes)
    ok = ok and got == [_alone(p, m)[0] for p, m in zip(prompts, maxes)]
    ret[rank] = ok
    dist.destroy_process_group()


def test_broadcast_shard_reduce_world2():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert all(ret[r] for r in range(world)), dict(ret)


def test_router_policies():
    from mlx_vlm_b200.parallel import least_loaded, shard_requests
    assert least_loaded([3, 1, 2, 1]) == 1
    assert shard_requests(5, 8, 6) == [] and shard_requests(64, 8, 3) == list(range(3, 64, 8))
