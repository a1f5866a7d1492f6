"""CPU (gloo, world_size 2): the host-side logic of the multi-GPU path — sample sharding and the flat-gradient mean
(reference: experiment.py:104-110, 159-160).  On the GPU box the same functions run over NCCL."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from view_fusion_b200.distributed import allreduce_mean_, broadcast_parameters, chunk_bounds, data_parallel, shard_samples


def test_shard_samples_equal_counts_is_reference_split():
    assert shard_samples([6] * 28, 1) == [(0, 28)]
    assert shard_samples([6] * 28, 2) == [(0, 14), (14, 28)]
    assert shard_samples([6] * 160, 8) == [(20 * r, 20 * r + 20) for r in range(8)]
    b = shard_samples([6] * 28, 8)                       # 28 = 8*3 + 4: contiguous, complete, sizes 3 or 4
    assert b[0][0] == 0 and b[-1][1] == 28 and all(b[i][1] == b[i + 1][0] for i in range(7))
    assert {e - s for s, e in b} <= {3, 4}


def test_shard_samples_ragged_balances_view_images():
    vc = [1, 6, 6, 1, 1, 1, 2, 6]                        # 24 view-images
    b = shard_samples(vc, 2)
    assert b[0][0] == 0 and b[0][1] == b[1][0] and b[1][1] == len(vc)
    loads = [sum(vc[s:e]) for s, e in b]
    assert max(loads) - min(loads) <= 6 and all(e > s for s, e in b)
    b4 = shard_samples(vc, 4)
    assert [s for s, _ in b4] == sorted(s for s, _ in b4) and all(e > s for s, e in b4) and b4[-1][1] == len(vc)
    with pytest.raises(ValueError):
        shard_samples([6, 6], 3)


def test_chunk_bounds_cover_exactly():
    for n, c in [(33_947_206, 4), (1000, 4), (5, 8), (4096, 1)]:
        b = chunk_bounds(n, c)
        assert b[0][0] == 0 and b[-1][1] == n and len(b) <= max(1, c)
        assert all(b[i][1] == b[i + 1][0] for i in range(len(b) - 1))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)
        n = 10_007
        flat = torch.randn(n)
        mine = flat.clone()
        allreduce_mean_(flat, None, chunks=4)
        gathered = [torch.empty(n) for _ in range(world)]
        dist.all_gather(gathered, mine)
        expect = torch.stack(gathered).mean(0)
        ok_mean = bool(torch.allclose(flat, expect, atol=1e-6))
        # parameter broadcast + the data_parallel switch on a stand-in module
        lin = torch.nn.Linear(7, 5)
        data_parallel(lin, None, chunks=2)
        w = [torch.empty_like(lin.weight) for _ in range(world)]
        dist.all_gather(w, lin.weight.data)
        ok_bcast = bool(all(torch.equal(w[0], x) for x in w)) and lin._grad_sync == (None, 2)
        # sharded "training": each rank's gradient on its shard, averaged == full-batch gradient (equal shard sizes)
        torch.manual_seed(7)
        X, Y = torch.randn(8, 7), torch.randn(8, 5)
        s, e = shard_samples([6] * 8, world)[rank]
        lin.zero_grad()
        torch.nn.functional.mse_loss(lin(X[s:e]), Y[s:e]).backward()
        g = torch.cat([lin.weight.grad.reshape(-1), lin.bias.grad.reshape(-1)])
        allreduce_mean_(g, None, chunks=3)
        ref = torch.nn.Linear(7, 5)
        ref.load_state_dict(lin.state_dict())
        torch.nn.functional.mse_loss(ref(X), Y).backward()
        gr = torch.cat([ref.weight.grad.reshape(-1), ref.bias.grad.reshape(-1)])
        ok_dp = bool(torch.allclose(g, gr, atol=1e-6))
        q.put((rank, ok_mean, ok_bcast, ok_dp))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gloo_world2_gradient_mean_and_broadcast():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] and r[3] for r in res), res


def test_single_process_is_a_no_op():
    x = torch.arange(10.0)
    assert allreduce_mean_(x.clone()).equal(x)
    broadcast_parameters(torch.nn.Linear(2, 2))
