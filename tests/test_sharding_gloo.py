"""N > 1 path on CPU: world_size-2 gloo run of the sharding + ordered gather (the compute is the oracle here;
on the GPU box each rank calls the C-ABI on its own device instead)."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from phylocsfpp_b200.shard import contiguous_partition, gather_ordered


def test_contiguous_partition_properties():
    for weights, parts in (([5, 1, 1, 1, 8, 2, 2], 3), ([10], 4), ([], 2), ([3] * 16, 8), ([1, 100, 1], 2)):
        ranges = contiguous_partition(weights, parts)
        assert len(ranges) == parts and ranges[0][0] == 0 and ranges[-1][1] == len(weights)
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    r = contiguous_partition([3] * 16, 8)
    assert all(hi - lo == 2 for lo, hi in r)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as orc
    from phylocsfpp_b200.models import load_model
    from tests.util import random_alignment
    model = load_model("7yeast")
    alns = [random_alignment(model.nl, L, seed=L) for L in (30, 9, 120, 45, 60, 15, 200)]
    lo, hi = contiguous_partition([a.shape[1] for a in alns], world)[rank]
    mc, mnc = orc.OracleModel(model.tree, model.S_c, model.f_c), orc.OracleModel(model.tree, model.S_nc, model.f_nc)
    local = [float(orc.run_fixed(mc, mnc, orc.translate(a), False)[0]) for a in alns[lo:hi]]
    parts = gather_ordered(local, dst=0)
    if rank == 0:
        merged = [x for p in parts for x in p]
        single = [float(orc.run_fixed(mc, mnc, orc.translate(a), False)[0]) for a in alns]
        ret["ok"] = merged == single and len(merged) == len(alns)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_ordered_gather_matches_single_process():
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
        assert ret.get("ok") is True
