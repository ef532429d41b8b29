"""Multi-GPU row-strip lattices: one process per GPU, torch.distributed only for plumbing.

The data path has no host-driven collective: each rank's fused step kernel pushes its boundary rows
(V, last_firing_time, neurotransmitter t) straight into the neighbouring GPUs' ghost slots over NVLink
peer memory and raises an arrival counter there; the neighbour's boundary warps spin on that counter
before they gather (csrc/kernels.cu).  torch.distributed moves the CUDA-IPC blobs once at set-up and
provides barriers around timing.
"""
from __future__ import annotations

from . import _capi as K
from .backend import CudaLatticeBackend


def partition_rows(rows: int, world: int, rank: int):
    """[begin, end) of `rank`'s strip — the library's own host-side plan (snn_partition_begin)."""
    lib = K.load_library()
    return lib.snn_partition_begin(rows, world, rank), lib.snn_partition_begin(rows, world, rank + 1)


def exchange_blobs(blob: bytes, rank: int, world: int, group=None):
    """all-gather of the opaque IPC blobs; returns (blob of rank-1 or None, blob of rank+1 or None)."""
    import torch.distributed as dist
    blobs = [None] * world
    dist.all_gather_object(blobs, blob, group=group)
    return (blobs[rank - 1] if rank > 0 else None), (blobs[rank + 1] if rank < world - 1 else None)


class StripLattice:
    """A lattice of `rows` x `cols` neurons split into contiguous row strips, one per rank."""

    def __init__(self, model, rows, cols, rank, world, ntk=K.NT_APPROXIMATE, rck=K.RC_APPROXIMATE, device=-1):
        self.rank, self.world = rank, world
        self.rows, self.cols = rows, cols
        self.row_begin, self.row_end = partition_rows(rows, world, rank)
        self.be = CudaLatticeBackend(model, ntk, rck, rows, cols, device=device, rank=rank, world=world)
        self.local_rows = self.row_end - self.row_begin
        self.n_local = self.local_rows * cols

    def attach(self, group=None):
        """Export my halo-visible slab, gather everyone's, map the two neighbours.  Call after the fields that
        decide the memory layout (chemistry on/off, stencil radius) have been set."""
        if self.world == 1:
            return
        import torch.distributed as dist
        lo, hi = exchange_blobs(self.be.ipc_export(), self.rank, self.world, group)
        if lo is not None:
            self.be.ipc_attach(-1, lo)
        if hi is not None:
            self.be.ipc_attach(+1, hi)
        dist.barrier(group=group)
