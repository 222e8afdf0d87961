"""world_size-2 gloo tests (CPU) of the data-parallel host logic: sharding, caption gather, the single gradient all-reduce."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import imagecaptioning.pytorch_b200 as b200
    P = b200.parallel
    n_total, T = 7, 5
    full = torch.arange(n_total * T).reshape(n_total, T)
    s, e = P.shard_range(n_total, rank, world)
    local = full[s:e].clone()
    gathered = P.gather_captions(local, n_total)
    ok_gather = torch.equal(gathered, full)
    # gradient averaging: rank r holds grad = (r + 1) * ones -> mean over ranks
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(3, 4)), torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(2))]
    params[0].grad = torch.full((3, 4), float(rank + 1))
    params[1].grad = torch.arange(5.0) * (rank + 1)
    nbytes = P.allreduce_gradients(params)          # params[2] has no grad and must be skipped
    mean = sum(range(1, world + 1)) / world
    ok_grad = torch.allclose(params[0].grad, torch.full((3, 4), mean)) and torch.allclose(params[1].grad, torch.arange(5.0) * mean) and params[2].grad is None
    mx = P.max_over_ranks(float(rank), 'cpu')
    q.put((rank, bool(ok_gather), bool(ok_grad), nbytes, mx, (s, e)))
    dist.destroy_process_group()


def test_two_rank_gloo():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert [r[5] for r in res] == [(0, 4), (4, 7)]
    for rank, ok_gather, ok_grad, nbytes, mx, _ in res:
        assert ok_gather and ok_grad and nbytes == (12 + 5) * 4 and mx == 1.0


def test_shard_range_covers_everything():
    import imagecaptioning.pytorch_b200 as b200
    for n in (0, 1, 7, 64, 513):
        for world in (1, 2, 3, 8):
            spans = [b200.parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(e - s for s, e in spans) - min(e - s for s, e in spans) <= 1
