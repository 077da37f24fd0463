"""Multi-GPU plumbing: one process per GPU, z-slab (last-axis) partition.  torch.distributed is
used only to ship the NCCL unique id and for barriers; the data path (halo exchange, dot-product
allreduce) runs inside libtopopt_cuda over its own NCCL communicator."""
from __future__ import annotations

import ctypes as C
import os

from . import _lib


class SlabComm:
    """rank/world of the slab decomposition.  Every handle needs its own NCCL communicator, hence
    its own ncclUniqueId: ``fresh_id()`` is a collective call (rank 0 draws, everyone receives)."""

    def __init__(self, rank, world):
        self.rank, self.world = rank, world

    def fresh_id(self):
        import torch
        import torch.distributed as dist

        buf = C.create_string_buffer(128)
        if self.rank == 0:
            _lib.check(_lib.load().topopt_nccl_unique_id(C.cast(buf, C.c_void_p)))
        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.tensor(list(buf.raw), dtype=torch.uint8, device=dev)
        dist.broadcast(t, src=0)
        return bytes(t.cpu().tolist())

    @property
    def nccl_id(self):
        return self.fresh_id()

    def all_gather_bytes(self, blob: bytes) -> bytes:
        """Concatenation of every rank's equally sized byte string, in rank order (collective)."""
        import torch
        import torch.distributed as dist

        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        mine = torch.tensor(list(blob), dtype=torch.uint8, device=dev)
        out = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(out, mine)
        return b"".join(bytes(t.cpu().tolist()) for t in out)

    def all_agree(self, ok: bool) -> bool:
        """True iff ``ok`` is true on every rank (collective)."""
        import torch
        import torch.distributed as dist

        dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
        t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(int(t.item()))


def slab_ranges(nlayers, world):
    """element layers [e0, e1) and owned node planes [k0, k1) per rank (the upper rank owns a
    shared plane; the last rank also owns the top plane)."""
    out = []
    for r in range(world):
        e0, e1 = r * nlayers // world, (r + 1) * nlayers // world
        k1 = nlayers + 1 if r == world - 1 else e1
        out.append(((e0, e1), (e0, k1)))
    return out


def init_from_env(backend=None):
    """Under torchrun: initialise torch.distributed, broadcast rank 0's ncclUniqueId, return
    (SlabComm | None, local device ordinal)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1:
        return None, local
    import torch
    import torch.distributed as dist

    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend)
    return SlabComm(rank, world), local
