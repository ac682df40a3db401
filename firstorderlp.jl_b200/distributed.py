"""One process per GPU: the torch.distributed plumbing around folp_dist.

The library itself only needs (rank, world_size, device, ncclUniqueId); this
module gets them from a torchrun environment and broadcasts the id rank 0
creates. After `init()`, every `Solver(...)` built without an explicit `dist`
joins a fresh NCCL communicator of all ranks -- so `folp_b200.optimize` runs
row-partitioned under torchrun with no change at the call site. All ranks must
construct their solvers (and call them) in the same order.
"""
from __future__ import annotations

import os
from typing import Optional

_STATE: Optional[dict] = None
_UID: Optional[bytes] = None  # the library keeps one NCCL communicator per process (and reuses it), so one id suffices


def init(backend: str = "nccl") -> Optional[dict]:
    """Reads RANK / WORLD_SIZE / LOCAL_RANK (torchrun). Returns None when world_size == 1."""
    global _STATE
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        _STATE = None
        return None
    import torch
    import torch.distributed as td

    rank = int(os.environ["RANK"])
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if backend == "nccl":
        torch.cuda.set_device(local)
    if not td.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend == "nccl":
            td.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            td.init_process_group(backend)
    _STATE = {"rank": rank, "world_size": world, "device": local, "backend": backend}
    return _STATE


def state() -> Optional[dict]:
    return _STATE


def new_dist():
    """Collective over all ranks: a folp_dist carrying a fresh ncclUniqueId (None when single)."""
    global _UID
    if _STATE is None:
        return None
    from . import lib

    if _UID is not None and not os.environ.get("FOLP_NO_COMM_CACHE"):
        return lib.make_dist(_STATE["rank"], _STATE["world_size"], _STATE["device"], _UID)
    import torch
    import torch.distributed as td

    dev = torch.device("cuda", _STATE["device"]) if _STATE["backend"] == "nccl" else torch.device("cpu")
    buf = torch.zeros(128, dtype=torch.uint8, device=dev)
    if _STATE["rank"] == 0:
        buf.copy_(torch.frombuffer(bytearray(lib.nccl_unique_id()), dtype=torch.uint8))
    td.broadcast(buf, src=0)
    uid = bytes(buf.cpu().numpy().tobytes())
    _UID = uid
    return lib.make_dist(_STATE["rank"], _STATE["world_size"], _STATE["device"], uid)
