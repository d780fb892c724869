"""Multi-process host logic on CPU: world_size 2 over gloo (the N>1 path of parallel.py / sampler.py)."""
import os
import socket
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from hierdiff_b200 import parallel
    ctx = parallel.init(backend="gloo")
    assert (ctx.rank, ctx.world) == (rank, world) and ctx.device.type == "cpu"
    # one broadcast of the flat parameter buffer: every rank ends with rank 0's weights
    torch.manual_seed(100 + rank)
    net = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3))
    net.register_buffer("buf", torch.full((2,), float(rank)))
    nbytes = parallel.broadcast_parameters(net, ctx)
    assert nbytes == 4 * (7 * 5 + 5 + 5 * 3 + 3 + 2)
    torch.manual_seed(100)
    want = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3))
    for a, b in zip(net.parameters(), want.parameters()):
        assert torch.equal(a, b)
    assert torch.equal(net.buf, torch.zeros(2))
    # batches are sharded in contiguous blocks, remainder to the low ranks
    assert parallel.shard_count(5, ctx) == (3 if rank == 0 else 2)
    assert parallel.shard_range(5, ctx) == ((0, 3) if rank == 0 else (3, 5))
    assert parallel.shard_count(1, ctx) == (1 if rank == 0 else 0)
    # results come back in rank order on rank 0 only, in the pickle layout of sampler.py:40-41
    lo, hi = parallel.shard_range(5, ctx)
    local = ([{"x": torch.full((2, 3), float(k)), "h": torch.zeros(2, 8)} for k in range(lo, hi)], [])
    merged = parallel.gather_results(local, ctx)
    if rank == 0:
        assert [int(r["x"][0, 0]) for r in merged[0]] == [0, 1, 2, 3, 4] and merged[1] == []
    else:
        assert merged is None
    # timing is the max over ranks
    assert parallel.max_over_ranks(float(rank + 1), ctx) == float(world)
    parallel.barrier(ctx)
    parallel.finish(ctx)
    open(os.path.join(out_dir, f"ok{rank}"), "w").close()


def test_two_rank_gloo(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def test_single_process_is_a_noop():
    sys.path.insert(0, ROOT)
    from hierdiff_b200 import parallel
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"):
        os.environ.pop(k, None)
    ctx = parallel.init()
    assert ctx.world == 1 and parallel.shard_range(7, ctx) == (0, 7)
    assert parallel.broadcast_parameters(torch.nn.Linear(2, 2), ctx) == 0
    assert parallel.gather_results(([1], [2]), ctx) == ([1], [2])
    assert parallel.max_over_ranks(3.5, ctx) == 3.5
