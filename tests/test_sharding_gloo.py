"""N > 1 path on CPU: world_size-2 gloo run of the sharding + ordered gather (the compute is the oracle here;
on the GPU box each rank calls the C-ABI on its own device instead)."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from phylocsfpp_b200.shard import OrderedHostBuffer, column_ranges, contiguous_partition, gather_ordered


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


def _worker_shared(rank, world, port, path, total):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    if rank == 0:
        OrderedHostBuffer(path, 2, total, create=True).close()
    dist.barrier()
    buf = OrderedHostBuffer(path, 2, total)
    lo, hi = column_ranges(total, world, align=16)[rank]
    for c0 in range(lo, hi, 1000):          # batches, as bench.py's config-4 leg writes them
        c1 = min(hi, c0 + 1000)
        buf.write(0, c0, np.arange(c0, c1, dtype=np.float64))
        buf.write(1, c0, -np.arange(c0, c1, dtype=np.float64))
    buf.arr.flush()
    dist.barrier()
    buf.close()
    dist.destroy_process_group()


def test_column_ranges_and_ordered_host_buffer(tmp_path):
    """Strong-scaled column ranges + the file-backed ordered output: two ranks write their ranges, the merged array is the
    single-process array (no collective on the data path)."""
    for total, parts, align in ((1000, 3, 16), (250_000_000, 8, 1 << 22), (10, 4, 16), (0, 2, 1)):
        r = column_ranges(total, parts, align)
        assert r[0][0] == 0 and r[-1][1] == total and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert all(lo % align == 0 or lo == total for lo, _ in r)
    total, path = 12345, os.path.join(str(tmp_path), "ordered.bin")
    mp.spawn(_worker_shared, args=(2, _free_port(), path, total), nprocs=2, join=True)
    merged = np.memmap(path, dtype=np.float64, mode="r", shape=(2, total))
    assert np.array_equal(merged[0], np.arange(total)) and np.array_equal(merged[1], -np.arange(total, dtype=np.float64))
