"""Multi-GPU sharding of the hot path: independent units, zero exchange.

The reference's only parallelism is an OpenMP loop over MAF byte ranges whose per-job outputs are merged in
job order (src/phylocsf++build_tracks.hpp:88, :27-53).  Here the unit is one concatenated alignment chain (or one
score-msa block); ranks get contiguous, column-balanced ranges so that concatenating the per-rank outputs in rank
order reproduces the single-process output byte for byte.  There is no data-path collective; the only
communication is the host-side ordered gather of the (small) per-rank results.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple


def contiguous_partition(weights: Sequence[int], parts: int) -> List[Tuple[int, int]]:
    """Splits items 0..n-1 into `parts` contiguous ranges [lo, hi) with near-equal total weight (greedy on the
    cumulative sum; a range may be empty when there are fewer items than parts)."""
    n = len(weights)
    total = sum(weights)
    out, lo, acc = [], 0, 0
    for p in range(parts):
        target = total * (p + 1) / parts
        hi = lo
        while hi < n and (acc + weights[hi] <= target or (hi == lo and p < parts - 1 and n - hi > parts - 1 - p)):
            acc += weights[hi]
            hi += 1
        if p == parts - 1:
            hi = n
        out.append((lo, hi))
        lo = hi
    return out


def gather_ordered(obj, dst: int = 0):
    """Host-side ordered gather: returns [obj_rank0, obj_rank1, ...] on `dst`, None elsewhere.
    Works with any initialised torch.distributed backend (gloo on CPU, nccl on GPUs); single process -> [obj]."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [obj]
    world, rank = dist.get_world_size(), dist.get_rank()
    if dist.get_backend() == "nccl":
        out = [None] * world
        dist.all_gather_object(out, obj)
        return out if rank == dst else None
    out = [None] * world if rank == dst else None
    dist.gather_object(obj, out, dst=dst)
    return out


def column_ranges(total: int, parts: int, align: int = 1) -> List[Tuple[int, int]]:
    """[lo, hi) column range of every rank for a strong-scaled run over `total` columns (boundaries on multiples of `align`)."""
    cuts = [min(total, (total * p // parts + align - 1) // align * align) for p in range(parts)] + [total]
    return [(cuts[p], max(cuts[p], cuts[p + 1])) for p in range(parts)]


class OrderedHostBuffer:
    """The host-side ordered gather of a one-node multi-process run without any collective: one [rows, total] array backed by a
    file that every rank maps (tmpfs when it has the room), each rank writing the columns of its own range.  When the ranks are
    done, rank 0 holds the merged tracks in column order — the analogue of the reference's merge of per-job files in job order
    (src/phylocsf++build_tracks.hpp:27-53, 245-259).  `path` must be the same on every rank (e.g. derived from MASTER_PORT)."""

    def __init__(self, path: str, rows: int, total: int, dtype="float64", create: bool = False):
        import numpy as np
        self.path, self.rows, self.total = path, rows, total
        self.dtype = np.dtype(dtype)
        nbytes = rows * total * self.dtype.itemsize
        if create:
            import os
            with open(path, "wb") as fh:
                fh.truncate(nbytes)
                if nbytes:
                    try:
                        os.posix_fallocate(fh.fileno(), 0, nbytes)          # the pages exist before anybody writes (tmpfs: no allocation faults later)
                    except OSError:
                        pass
        self.arr = np.memmap(path, dtype=self.dtype, mode="r+", shape=(rows, total))

    def prefault(self, col0: int, col1: int) -> None:
        """Touches this rank's column range once (page-table entries of the shared mapping), as one would for any output buffer."""
        if col1 > col0:
            step = max(1, 4096 // self.dtype.itemsize)
            self.arr[:, col0:col1:step] = 0          # a write, so that the pages are mapped writable (the file is all zeros anyway)

    @staticmethod
    def pick_dir(nbytes: int) -> str:
        """tmpfs if it has room for nbytes (+10 %), else the system temporary directory."""
        import os
        import shutil
        import tempfile
        try:
            if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > nbytes * 1.1:
                return "/dev/shm"
        except OSError:
            pass
        return tempfile.gettempdir()

    def write(self, row: int, col0: int, values) -> None:
        self.arr[row, col0:col0 + len(values)] = values

    def close(self, unlink: bool = False) -> None:
        import os
        del self.arr
        if unlink:
            try:
                os.unlink(self.path)
            except OSError:
                pass
