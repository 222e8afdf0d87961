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
    # the flat gradient buffer of the fused steps: ONE all-reduce over the whole buffer averages every view in place
    from imagecaptioning.pytorch_b200.grad_sync import FlatGrads, GradSync, allreduce_flat
    fg = FlatGrads([[params[0]], [params[1], params[2]]], 'cpu')
    fg.view(params[0]).fill_(float(rank + 1))
    fg.view(params[1]).copy_(torch.arange(5.0) * (rank + 1))
    fg.view(params[2]).fill_(10.0 * rank)
    flat_bytes = allreduce_flat(fg)
    ok_flat = (torch.allclose(fg.view(params[0]), torch.full((3, 4), mean)) and torch.allclose(fg.view(params[1]), torch.arange(5.0) * mean)
               and torch.allclose(fg.view(params[2]), torch.full((2,), 10.0 * (world - 1) / 2)) and flat_bytes == fg.nbytes and GradSync().world == world)
    q.put((rank, bool(ok_gather), bool(ok_grad and ok_flat), nbytes, mx, (s, e)))
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


def test_flat_gradient_buffer_layout():
    """grad_sync.FlatGrads: every view starts on a 256-byte boundary, views are disjoint, groups are contiguous ranges in the order given
    (the order the engine completes them), and the ranges tile the buffer -- what the engines' pointer tables and the per-group all-reduce
    chunks rely on."""
    from imagecaptioning.pytorch_b200.grad_sync import ALIGN, FlatGrads
    shapes = [[(7, 3), (5,)], [(1,)], [(64,), (65,), (2, 2, 2)]]
    groups = [[torch.nn.Parameter(torch.zeros(*s)) for s in g] for g in shapes]
    fg = FlatGrads(groups, 'cpu')
    assert fg.event_table()[1] == 0 and fg.events == []                 # events exist on CUDA devices only
    base = fg.flat.data_ptr()
    at = 0
    for g, (s, e) in zip(groups, fg.ranges):
        assert s == at
        for p in g:
            v = fg.view(p)
            assert v.shape == p.shape and v.is_contiguous()
            off = (v.data_ptr() - base) // 4
            assert off == at and off % ALIGN == 0 and (v.data_ptr() - base) % 256 == 0
            at += (p.numel() + ALIGN - 1) // ALIGN * ALIGN
        assert e == at
    assert at == fg.flat.numel() and fg.nbytes == at * 4
    assert [id(p) for p in fg.params] == [id(p) for g in groups for p in g]
    # writing one view leaves every other element of the buffer untouched
    fg.view(groups[2][1]).fill_(1.0)
    assert float(fg.flat.sum()) == 65.0
