"""One whole-graph SSPPR query split over the GPUs of a box (BASELINE.json config 5).

Push runs on rank 0; its (reserve, residue) state is broadcast into every rank's engine buffers; each rank
walks its share of the SAME walk plan (Philox keyed by source / walk index, independent of the split) into a
dense fp64 vector; the vectors are summed with one all-reduce (NCCL over NVLink / NVSwitch).  torch.distributed
is plumbing only: it moves the two vectors, all arithmetic happens in libfora_b200.so.
"""
import numpy as np


class _DevArray:
    """expose a raw device pointer to torch through __cuda_array_interface__ (no copy)"""

    def __init__(self, ptr, n, dtype="<f8"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": dtype, "data": (int(ptr), False), "version": 3, "strides": None}


def device_tensor(ptr, n, device):
    import torch
    return torch.as_tensor(_DevArray(ptr, n), device=device)


def to_original(engine, internal_tensor):
    """dense vector in the engine's internal vertex order -> a new CUDA tensor indexed by original vertex id"""
    import torch
    out = torch.empty_like(internal_tensor)
    engine.device_to_original(internal_tensor.data_ptr(), out.data_ptr())
    return out


def ssppr_split(engine, source, rmax, qid=0, group=None):
    """returns (ppr as a torch CUDA tensor indexed by original vertex id, stats).  Call on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    n = engine.n
    dev = torch.device("cuda", torch.cuda.current_device())
    engine.prepare_slots()
    reserve = device_tensor(engine.device_reserve_ptr(0), n, dev)
    residue = device_tensor(engine.device_residue_ptr(0), n, dev)
    meta = torch.zeros(1, dtype=torch.float64, device=dev)
    if rank == 0:
        res_h, rsd_h, rsum, st = engine.push_only(int(source), rmax)  # leaves the state in slot 0 on rank 0's GPU
        meta[0] = rsum
    torch.cuda.synchronize()
    if world > 1:
        dist.broadcast(reserve, 0, group=group)
        dist.broadcast(residue, 0, group=group)
        dist.broadcast(meta, 0, group=group)
    torch.cuda.synchronize()
    stats = engine.compute_ppr_part_device(float(meta.item()), qid, rank, world)
    if world > 1:
        dist.all_reduce(reserve, op=dist.ReduceOp.SUM, group=group)  # the PPR vector now lives in `reserve` (in place)
    torch.cuda.synchronize()
    return to_original(engine, reserve), stats
