"""CPU, world_size 2, gloo: the host-side logic of the data-parallel path -- batch partition, flat gradient
arena, one all-reduce per network, identical replicas afterwards, metric averaging."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ipr_gan_b200 import dist as ipr_dist, flat
        torch.manual_seed(0)                                   # identical replicas
        net = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 3))
        arena = flat.arena_for(list(net.parameters()))
        assert all(p.data_ptr() >= arena.param.data_ptr() for p in net.parameters())
        g = torch.Generator().manual_seed(1)
        x, y = torch.randn(8, 16, generator=g), torch.randn(8, 3, generator=g)
        xs, ys = ipr_dist.shard(x, rank, world), ipr_dist.shard(y, rank, world)
        assert xs.shape[0] == 4 and torch.equal(xs, x[rank * 4:(rank + 1) * 4])
        arena.zero_grad()
        torch.nn.functional.mse_loss(net(xs), ys).backward()    # accumulates into the arena views
        assert float(arena.grad.abs().sum()) > 0
        ipr_dist.allreduce_sum_(arena.reduce_view)               # what FlatAdam.step() issues; 1/world rides in Adam
        arena.grad.mul_(1.0 / world)
        # reference: the full batch on one process
        torch.manual_seed(0)
        ref = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh(), torch.nn.Linear(32, 3))
        torch.nn.functional.mse_loss(ref(x), y).backward()
        for p, q in zip(net.parameters(), ref.parameters()):
            assert torch.allclose(p.grad, q.grad, rtol=1e-5, atol=1e-7)
        # replicas hold identical reduced gradients
        gathered = [torch.empty_like(arena.grad) for _ in range(world)]
        dist.all_gather(gathered, arena.grad)
        assert torch.equal(gathered[0], gathered[1])
        m = ipr_dist.reduce_metrics({"a": float(rank), "b": 2.0})
        assert m == {"a": 0.5, "b": 2.0}
        out[rank] = 1
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_allreduce_gloo():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert dict(out) == {0: 1, 1: 1}


def test_arena_views_and_zero_grad():
    from ipr_gan_b200 import flat
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 3))
    before = [p.detach().clone() for p in net.parameters()]
    arena = flat.arena_for(list(net.parameters()))
    for p, b in zip(net.parameters(), before):
        assert torch.equal(p, b) and p.data_ptr() % 16 == 0
    assert flat.arena_for(list(net.parameters())) is arena               # idempotent
    net(torch.randn(2, 5)).sum().backward()
    assert float(arena.grad.abs().sum()) > 0
    arena.zero_grad()
    assert float(arena.grad.abs().sum()) == 0 and all(p.grad is not None for p in net.parameters())
    sd = net.state_dict()
    assert list(sd.keys()) == ["0.weight", "0.bias", "1.weight", "1.bias"]


def test_shared_gradient_buffer_layout():
    """[D grads | D slots | G slots | G grads]: each network's all-reduce range covers its gradients and its own four
    metric slots and nothing of the other network; the eight slots are one contiguous board."""
    from ipr_gan_b200 import flat
    d = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.Linear(7, 1))
    g = torch.nn.Sequential(torch.nn.Linear(3, 6), torch.nn.Linear(6, 5))
    ad, ag = flat.arena_for(list(d.parameters())), flat.arena_for(list(g.parameters()))
    d(torch.randn(2, 5)).sum().backward()
    before = [p.grad.clone() for p in d.parameters()]
    board = flat.share_gradient_buffer(ad, ag)
    assert board.numel() == 8 and board.data_ptr() == ad.grad.data_ptr() + 4 * ad.numel
    assert ag.grad.data_ptr() == board.data_ptr() + 32 and ag.grad.data_ptr() % 16 == 0
    assert ad.reduce_view.data_ptr() == ad.grad.data_ptr() and ad.reduce_view.numel() == ad.numel + 4
    assert ag.reduce_view.data_ptr() == ag.slots.data_ptr() and ag.reduce_view.numel() == ag.numel + 4
    assert ad.slots.data_ptr() == board.data_ptr() and ag.slots.data_ptr() == board.data_ptr() + 16
    for p, b in zip(d.parameters(), before):                       # gradients survived the move, views re-bound
        assert torch.equal(p.grad, b) and ad.grad.data_ptr() <= p.grad.data_ptr() < board.data_ptr()
    board.fill_(3.0)
    ad.zero_grad(), ag.zero_grad()
    assert float(board.sum()) == 24.0 and float(ad.grad.abs().sum()) == 0     # zero_grad leaves the board alone
    g(torch.randn(2, 3)).sum().backward()
    assert float(ag.grad.abs().sum()) > 0 and float(board.sum()) == 24.0
