"""Multi-GPU plumbing for the augmentation path.

The path shards embarrassingly by sample: one process per GPU, rank r owns a contiguous slice of
the global batch, and there is NO data-path collective.  The only exchanges are a tiny control block
{seed, epoch} broadcast from rank 0 once per epoch (and, optionally, the generator's mixing weights when G runs on one
rank: broadcast_mix_weights), so every rank derives the same Philox keys
(seed, global sample index, op) and the result is independent of the number of ranks.
The reference's equivalent is `torch.nn.DataParallel` scatter (tools/train.py:69,106) after CPU
DataLoader workers; here each rank augments its own shard on its own GPU.
"""
import torch
import torch.distributed as dist


def shard_range(n_global, rank, world_size):
    """Contiguous [lo, hi) slice of a global batch of n_global samples (sizes differ by at most 1)."""
    base, rem = divmod(n_global, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def broadcast_control(seed, epoch, device=None, src=0):
    """Rank `src` decides (seed, epoch); everybody returns the same pair.  NCCL on GPUs, gloo on CPU."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return int(seed), int(epoch)
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([int(seed), int(epoch)], dtype=torch.int64, device=device)
    dist.broadcast(t, src=src)
    return int(t[0].item()), int(t[1].item())


def broadcast_mix_weights(logits_or_weights, global_batch, rank=None, world_size=None, src=0):
    """The "generator on rank 0" arrangement of SURVEY 8e / north_star: rank `src` runs model_G on the whole global batch and
    broadcasts its `[B_global, K, H, W]` mixing logits (or weights) over NCCL (gloo on CPU); every rank returns ITS shard
    `[hi - lo, K, H, W]` (a view into the received buffer) for `chain_mix_from_logits` / `mix`.
    `logits_or_weights`: the full tensor on rank `src`; on the other ranks a tensor of the same shape / dtype to receive into
    (or a `(shape, dtype, device)` tuple, then the buffer is allocated here).  With one rank it just slices."""
    t = logits_or_weights
    if isinstance(t, tuple):
        shape, dtype, device = t
        t = torch.empty(shape, dtype=dtype, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        rank = dist.get_rank() if rank is None else rank
        world_size = dist.get_world_size() if world_size is None else world_size
        t = t.contiguous()
        dist.broadcast(t, src=src)
    else:
        rank, world_size = (0 if rank is None else rank), (1 if world_size is None else world_size)
    lo, hi = shard_range(int(global_batch), rank, world_size)
    return t[lo:hi]


def sample_base(epoch, step, global_batch, lo):
    """Global sample index of the first sample of this rank's shard: the Philox `sample_base` that makes
    random draws a function of (seed, epoch, step, global position) only."""
    return (int(epoch) * (1 << 32)) + int(step) * int(global_batch) + int(lo)


def bind_to_gpu_numa(device_index):
    """Pin this process to the CPUs that are local to GPU `device_index` (sysfs `local_cpulist` of its PCI function), so
    the pinned host buffers it allocates afterwards live on the GPU's own NUMA node.  With one process per GPU and every
    rank streaming ~48 GB/s out of pinned host memory, remote-node buffers saturate the inter-socket link long before
    PCIe.  Best effort: returns the CPU list it bound to, or None when the topology cannot be read (nothing changes)."""
    import os
    try:
        bus = torch.cuda.get_device_properties(device_index).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(device_index), "pci_domain_id", 0)
        dev = getattr(torch.cuda.get_device_properties(device_index), "pci_device_id", 0)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (dom, bus, dev)
        with open(path) as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None
