"""One process per GPU: contiguous sharding of the parameter vector + communicator bootstrap.

The optimizer state shards by contiguous blocks: rank r owns
[offset, offset + n_local) of every n-vector; inside a step the only exchange is one small
sum all-reduce of fp64 partial dots per reduction phase, done by the CUDA library itself on
its own stream (NCCL over NVLink).  ``torch.distributed`` is only the plumbing that ships the
128-byte communicator id from rank 0 to the others.
"""
from __future__ import annotations

import ctypes as C


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous block of rank `rank`: the first n % world ranks get one extra element.
    Returns (offset, n_local)."""
    base, extra = divmod(int(n), int(world))
    n_local = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, n_local


def broadcast_id(id_bytes, rank: int, src: int = 0, group=None) -> bytes:
    """Ship the 128-byte communicator id from `src` to every rank over torch.distributed
    (works on gloo and nccl groups: object broadcast)."""
    import torch.distributed as dist

    box = [bytes(id_bytes) if rank == src else None]
    dist.broadcast_object_list(box, src=src, group=group)
    return box[0]


def init_comm(abi, rank: int, world: int, group=None):
    """Create the library's communicator on every rank.  Returns an opaque handle (c_void_p)."""
    lib = abi.lib
    buf = (C.c_char * 128)()
    if rank == 0:
        if lib.stochqn_b200_comm_unique_id(buf) != 0:
            raise RuntimeError("comm_unique_id failed")
    raw = broadcast_id(bytes(buf.raw), rank, 0, group)
    buf2 = (C.c_char * 128).from_buffer_copy(raw)
    comm = C.c_void_p()
    if lib.stochqn_b200_comm_init(buf2, rank, world, C.byref(comm)) != 0:
        raise RuntimeError("comm_init failed on rank %d" % rank)
    return comm
