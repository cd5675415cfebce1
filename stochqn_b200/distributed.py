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


class RowShardedCombiner:
    """The two collectives of a ROW-sharded callback (BASELINE config 5's multi-GPU layout).

    Every rank evaluates the gradient (or Hessian-vector product) on ITS rows of the batch at the full point and the
    optimizer state is sharded by blocks, so each request needs

    * ``gather(block_ptr)``: the requested point, of which every rank holds block ``rank``, on every rank, and
    * ``reduce_scatter(out_block_ptr)``: block ``rank`` of the sum over ranks of the full-length partial vectors the
      callbacks wrote into ``send_buffer()``.

    With a peer-memory communicator both run as the library's own kernels over NVLink (push all-gather, pull
    reduce-scatter: the callback writes straight into the peer-mapped send vector, there is no staging copy) and the
    buffers belong to the library; otherwise they are ``ncclAllGather`` / ``ncclReduceScatter`` on two vectors this object
    allocates with torch.  Pointers are plain device addresses (ints), valid in stream order until the next call of the
    same method.  ``n`` must divide by ``world``; all ranks must make the same calls in the same order.
    """

    def __init__(self, abi, comm, n: int, world: int, stream=None, use_p2p=None):
        if n % world:
            raise ValueError("n = %d does not divide by the number of ranks (%d)" % (n, world))
        self.abi, self.lib, self.comm = abi, abi.lib, comm
        self.n, self.world, self.blk = int(n), int(world), int(n) // int(world)
        self.stream = stream
        has_p2p = bool(world > 1 and self.lib.stochqn_b200_comm_uses_p2p(comm))
        self.p2p = has_p2p if use_p2p is None else bool(use_p2p)
        if self.p2p and not has_p2p:
            raise RuntimeError("this communicator has no peer-memory path")
        self._send = None
        self._own = None
        if not self.p2p and self.world > 1:
            import torch
            tdt = torch.float64 if C.sizeof(abi.real) == 8 else torch.float32
            self._own = (torch.zeros(self.n, device="cuda", dtype=tdt), torch.zeros(self.n, device="cuda", dtype=tdt))

    def _check(self, rc, what):
        if rc != 0:
            from . import _lib
            raise RuntimeError("%s failed (%d): %s" % (what, rc, _lib.last_error(self.abi)))

    def gather(self, block_ptr: int) -> int:
        if self.world == 1:
            return int(block_ptr)
        if self.p2p:
            out = C.c_void_p()
            self._check(self.lib.stochqn_b200_all_gather_p2p(self.comm, block_ptr, self.blk, C.byref(out), self.stream), "all_gather_p2p")
            return int(out.value)
        self._check(self.lib.stochqn_b200_all_gather_real(self.comm, block_ptr, self._own[0].data_ptr(), self.blk, self.stream), "all_gather_real")
        return int(self._own[0].data_ptr())

    def send_buffer(self) -> int:
        """where the callback of THIS request writes its full-length partial vector (ask again for every request)"""
        if self.world == 1:
            raise RuntimeError("one rank: let the callback write into the optimizer's gradient directly")
        if self.p2p:
            out = C.c_void_p()
            self._check(self.lib.stochqn_b200_p2p_send_buffer(self.comm, self.blk, C.byref(out)), "p2p_send_buffer")
            self._send = int(out.value)
        else:
            self._send = int(self._own[1].data_ptr())
        return self._send

    def reduce_scatter(self, out_block_ptr: int) -> None:
        if self._send is None:
            raise RuntimeError("call send_buffer() and let the callback fill it first")
        if self.p2p:
            self._check(self.lib.stochqn_b200_reduce_scatter_p2p(self.comm, self._send, out_block_ptr, self.blk, self.stream), "reduce_scatter_p2p")
        else:
            self._check(self.lib.stochqn_b200_reduce_scatter_real(self.comm, self._send, out_block_ptr, self.blk, self.stream), "reduce_scatter_real")
        self._send = None
