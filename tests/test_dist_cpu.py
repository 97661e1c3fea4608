"""CPU test of the N>1 replica path: two gloo ranks on 127.0.0.1 shard an image list disjointly/exhaustively and
agree on the max-over-ranks timing and the summed throughput, as bench.py --gpus N reports them."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    from yolov5_obb_b200.dist_util import shard_range, max_over_ranks, sum_over_ranks, world
    assert world() == (rank, ws)
    mine = list(shard_range(37, rank, ws))
    dist.barrier()
    t = max_over_ranks(10.0 + rank)          # rank 1 is slower
    n = sum_over_ranks(float(len(mine)))
    q.put((rank, mine, t, n))
    dist.destroy_process_group()


def test_two_rank_replica_path():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, s0, t0, n0), (r1, s1, t1, n1) = res
    assert sorted(s0 + s1) == list(range(37)) and not set(s0) & set(s1) and abs(len(s0) - len(s1)) <= 1
    assert t0 == t1 == 11.0 and n0 == n1 == 37.0


def test_single_process_is_identity():
    from yolov5_obb_b200.dist_util import shard_range, max_over_ranks
    assert list(shard_range(5, 0, 1)) == [0, 1, 2, 3, 4] and max_over_ranks(3.5) == 3.5


def _grad_worker(rank, ws, port, q):
    """The training exchange step (train_step.TrainStep._allreduce_grads): one all-reduce of the flat gradient buffer
    gives every rank the SUM of the per-rank gradients (DDP mean of a loss x WORLD_SIZE, train.py:328)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    from types import SimpleNamespace
    from yolov5_obb_b200.train_step import TrainStep
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.BatchNorm2d(4), torch.nn.Conv2d(4, 2, 1))
    n = sum(p.numel() for p in net.parameters())
    flat = torch.arange(n, dtype=torch.float32) * (rank + 1)      # rank r holds (r+1) * [0, 1, 2, ...]
    o = 0
    for p in net.parameters():
        p.grad = flat[o:o + p.numel()].view(p.shape)
        o += p.numel()
    net._last_train_engine = SimpleNamespace(last_grad_flat=flat)
    ts = TrainStep.__new__(TrainStep)
    ts.model, ts.world = net, ws
    calls = []
    orig = dist.all_reduce
    dist.all_reduce = lambda t, op=dist.ReduceOp.SUM: (calls.append(t.numel()), orig(t, op=op))[1]
    ts._allreduce_grads()
    flat_calls = list(calls)
    # gradient accumulation: .grad still views the first step's buffer while the engine already holds a newer one
    net._last_train_engine.last_grad_flat = torch.zeros(n)
    calls.clear()
    ts._allreduce_grads()
    flat_calls += list(calls)
    # gradients that are NOT views of the flat buffer fall back to one call per tensor
    for p in net.parameters():
        p.grad = p.grad.clone()
    calls.clear()
    ts._allreduce_grads()
    dist.all_reduce = orig
    q.put((rank, flat.tolist(), flat_calls, len(calls), [p.grad.flatten()[:1].item() for p in net.parameters()][-1]))
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_is_one_call_and_a_sum():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    n = len(res[0][1])
    for rank, flat, flat_calls, per_tensor_calls, last in res:
        assert flat == [6.0 * i for i in range(n)]           # (1x + 2x) = 3x on both ranks, summed once more through ._base
        assert flat_calls == [n, n]                          # ONE collective over the whole buffer, both times
        assert per_tensor_calls == 6                         # fallback path: conv w/b, bn w/b, conv w/b


def _bcast_worker(rank, ws, port, q):
    """Replicas seeded per rank (init_seeds(1 + RANK), train.py:100) must start from rank 0's weights and BatchNorm
    statistics, as DDP's constructor broadcast guarantees (train.py:214): dist_util.broadcast_model_state, called by
    TrainStep.__init__ before the optimizer and the EMA copy are built."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(ws))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    from yolov5_obb_b200.dist_util import broadcast_model_state
    torch.manual_seed(1 + rank)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.BatchNorm2d(4), torch.nn.Conv2d(4, 2, 1))
    with torch.no_grad():
        net[1].running_mean += rank + 0.5
        net[1].num_batches_tracked += 7 * (rank + 1)
    ptrs = [p.data_ptr() for p in net.parameters()]
    before = torch.cat([p.detach().flatten() for p in net.parameters()]).clone()
    n = broadcast_model_state(net, 0)
    after = torch.cat([p.detach().flatten() for p in net.parameters()])
    q.put((rank, n, before.tolist(), after.tolist(), net[1].running_mean.tolist(), int(net[1].num_batches_tracked),
           ptrs == [p.data_ptr() for p in net.parameters()]))
    dist.destroy_process_group()


def test_two_rank_start_up_broadcast():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bcast_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (_, n0, b0, a0, rm0, nbt0, same0), (_, n1, b1, a1, rm1, nbt1, same1) = res
    assert b0 != b1                                   # the per-rank seeds really differ
    assert a0 == b0 and a1 == b0                      # every rank now holds rank 0's parameters
    assert rm0 == rm1 == [0.5] * 4 and nbt0 == nbt1 == 7
    assert n0 == n1 == 6 + 3 and same0 and same1      # 6 parameters + 3 buffers, written in place (plans keep their pointers)
