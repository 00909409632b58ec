"""Multi-GPU partition of a batch of independent streams: one process per GPU, contiguous
stream ranges, NO data-path collective — streams share no state, the read-only tables (sinc table, index
plan) are replicated per GPU.  torch.distributed is used only to agree on timing
(barrier + max over ranks) and, when a caller wants the results in one place, to gather them."""


def stream_range(n_streams, world, rank):
    """streams [lo, hi) of rank `rank`: the contiguous split [g*S/G, (g+1)*S/G) of SURVEY.md 8e"""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank {rank} of world {world}")
    return (rank * n_streams) // world, ((rank + 1) * n_streams) // world


def stream_seed(base_seed, config, stream):
    """seed of the synthetic generator for one global stream index: independent of how the
    batch is partitioned, so N-GPU runs process exactly the streams a 1-GPU run would"""
    return base_seed + (config << 20) + stream


def max_over_ranks(value, dist=None, device="cpu"):
    """max of a per-rank scalar (elapsed time) over all ranks; identity without a process group"""
    if dist is None or not dist.is_initialized():
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_counts(count, dist=None, device="cpu"):
    """units processed by every rank (list of length world); used for the whole-job throughput"""
    if dist is None or not dist.is_initialized():
        return [int(count)]
    import torch
    t = torch.tensor([int(count)], dtype=torch.int64, device=device)
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [int(x.item()) for x in out]


def bind_to_gpu_numa_node(device_index):
    """Pin this process (and the pinned host buffers it allocates afterwards, by first touch) to the NUMA node the GPU
    hangs off, so that N ranks staging host buffers do not all pull from one socket's memory.  Best effort: returns
    the node number, or None when the topology cannot be read (then nothing is changed)."""
    import os
    try:
        import torch
        props = torch.cuda.get_device_properties(device_index)
        bus = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return node
    except Exception:
        return None
