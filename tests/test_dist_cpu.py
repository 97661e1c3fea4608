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
