"""Batch sharding across GPUs (one process per GPU).  The reference has no multi-GPU code; images are
independent through conv, decode and NMS (utils.py:152-162 loops per image), so the path shards by
batch with exactly two collectives, both issued by libyolo_b200.so over NCCL:

  * once:      ncclBroadcast of the packed weight blob from rank 0     (yb_bcast_weights)
  * per batch: ncclAllGather of the fixed-capacity detection rows and  (yb_allgather_dets)
               their counts, enqueued on the same stream right after the NMS kernels

torch.distributed is used only as plumbing (shipping the 128-byte NCCL unique id, barriers).
The pure functions below hold the host-side logic and are what the gloo CPU tests exercise.
"""
from __future__ import annotations

import ctypes
from typing import List, Sequence, Tuple

import torch

from . import _lib


def shard_bounds(global_batch: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous image range [lo, hi) owned by `rank`; the first (global_batch % world) ranks get one extra."""
    if world <= 0 or not 0 <= rank < world or global_batch < 0:
        raise ValueError("bad shard request")
    base, extra = divmod(global_batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def merge_gathered(all_rows: torch.Tensor, all_counts: torch.Tensor, cand_any: bool = True) -> List[torch.Tensor]:
    """[G*B,cap,7] rows + [G*B] counts (rank-major = global image order for contiguous shards) ->
    per-image list in the reference's postprocessing() convention."""
    rows = all_rows.cpu()
    counts = all_counts.cpu().tolist()
    if not cand_any and sum(counts) == 0:
        return []
    return [rows[i, :c].clone() if c else torch.Tensor() for i, c in enumerate(counts)]


def exchange_unique_id(make_id, rank: int, group=None) -> bytes:
    """Rank 0 creates the NCCL unique id (make_id() -> 128 bytes); everyone receives it through
    torch.distributed (works on any backend, e.g. gloo in the CPU tests)."""
    import torch.distributed as dist
    box = [make_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    uid = box[0]
    if not isinstance(uid, (bytes, bytearray)) or len(uid) != 128:
        raise RuntimeError("unique id exchange failed")
    return bytes(uid)


class DetectionGather:
    """Owns the library-side NCCL communicator of one rank's YoloNet engine."""

    def __init__(self, net, rank: int, world: int, batch_local: int, cap: int):
        self.net, self.rank, self.world, self.B, self.cap = net, rank, world, batch_local, cap
        self.lib = _lib.load()
        self._inited = False
        self.all_rows = None
        self.all_counts = None

    def _init(self):
        if self._inited:
            return
        lib = self.lib

        def make_id():
            buf = (ctypes.c_uint8 * 128)()
            _lib.check(lib.yb_comm_unique_id(buf), None)
            return bytes(buf)

        uid = exchange_unique_id(make_id, self.rank)
        arr = (ctypes.c_uint8 * 128)(*uid)
        ctx = self.net._ctx
        if ctx is None:
            raise RuntimeError("run one forward first so that the engine exists")
        _lib.check(lib.yb_comm_init(ctx, arr, self.rank, self.world), ctx)
        dev = torch.device("cuda", self.net._ctx_device)
        self.all_rows = torch.empty(self.world * self.B, self.cap, 7, device=dev)
        self.all_counts = torch.empty(self.world * self.B, dtype=torch.int32, device=dev)
        self._inited = True

    def broadcast_weights(self, root: int = 0):
        self._init()
        ctx = self.net._ctx
        s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(self.lib.yb_bcast_weights(ctx, root, s), ctx)

    def allgather(self, rows: torch.Tensor, counts: torch.Tensor):
        self._init()
        ctx = self.net._ctx
        s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        _lib.check(self.lib.yb_allgather_dets(ctx, ctypes.c_void_p(rows.data_ptr()), ctypes.c_void_p(counts.data_ptr()),
                                              self.B, self.cap, ctypes.c_void_p(self.all_rows.data_ptr()),
                                              ctypes.c_void_p(self.all_counts.data_ptr()), s), ctx)
        return self.all_rows, self.all_counts
