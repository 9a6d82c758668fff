"""Host-side partitioning for the multi-GPU modes: queries are independent, the
graph is replicated per GPU, so there is no data-path collective -- only the partitioning below and one
all_reduce(MAX) of the elapsed time for reporting."""
import numpy as np


def query_block(n_queries, rank, world):
    """contiguous block [lo, hi) of the query list for `rank` (what ./fora --gpus N uses)"""
    return n_queries * rank // world, n_queries * (rank + 1) // world


def step_query_ids(queries, step, batch, rank, world):
    """weak scaling: every rank processes `batch` queries per step -- its own slice of a world*batch global
    batch cut from the (cyclic) query list.  Returns (ids, global query indices used as Philox keys)."""
    n = len(queries)
    base = step * batch * world + rank * batch
    idx = (base + np.arange(batch)) % n
    return np.asarray(queries)[idx].astype(np.int32), idx.astype(np.int64)


def balanced_source_ranges(offsets, counts, world):
    """index build sharded by source range, balanced by the number of walks (not by node count):
    returns world+1 cut points over [0, n]."""
    n = len(offsets)
    total = int(offsets[-1] + counts[-1]) if n else 0
    cuts = [0]
    for d in range(1, world):
        target = total * d // world
        cuts.append(int(np.searchsorted(offsets, target, side="left")))
    cuts.append(n)
    for i in range(1, len(cuts)):
        cuts[i] = max(cuts[i], cuts[i - 1])
    return cuts


def max_over_ranks(value):
    """elapsed time is reported as the max over ranks (device timers are per rank)"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-")
            cpus.update(range(int(a), int(b) + 1))
        else:
            cpus.add(int(part))
    return cpus


def bind_to_gpu_numa_node(device_index):
    """Pin this process to the CPUs of the NUMA node the GPU hangs off, so that pinned result buffers are
    allocated next to the GPU's PCIe root (the end-to-end path moves 38.8 MB per LJ-shape query).  Returns the
    node id, or None when the topology is not exposed (then nothing is changed)."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:  # nvml prints an 8-digit PCI domain, sysfs a 4-digit one
            bus = bus[4:]
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open("/sys/devices/system/node/node%d/cpulist" % node) as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None
