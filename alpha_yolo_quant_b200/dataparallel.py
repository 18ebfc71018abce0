"""Data-parallel host logic: shard a batch of images over the GPUs of one box, gather the detections in image order.

The hot path has no cross-image state (per-image input scale, per-image NMS), so ranks never exchange activations
(SURVEY.md 8(e)): every rank runs its own engine on a contiguous shard `ceil(N / G)` of the batch; only the results
(counts[N], dets[N,300,6]) and, for calibration, 64 per-tap maxima travel.  The collectives below run on whatever
backend the process group uses (NCCL tensors on the GPU box, gloo tensors in the CPU tests).
"""
import torch
import torch.distributed as dist

MAX_DET, DET_STRIDE = 300, 6


def shard_range(n_images, world_size, rank):
    """Contiguous shard [lo, hi) of rank `rank`: ceil(N / G) images per rank, the last ranks may be short or empty."""
    per = (n_images + world_size - 1) // world_size
    lo = min(rank * per, n_images)
    hi = min(lo + per, n_images)
    return lo, hi


def gather_detections(dets, counts, n_images, group=None, dst=0):
    """dets (m,300,6) / counts (m) of this rank's shard -> on rank `dst`: (N,300,6), (N) in image order; None elsewhere.
    Shards are padded to the common size ceil(N / G) so that one all_gather suffices."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = (n_images + world - 1) // world
    lo, hi = shard_range(n_images, world, rank)
    m = hi - lo
    assert dets.shape[0] == m and counts.shape[0] == m, (dets.shape, counts.shape, m)
    pd = torch.zeros((per, MAX_DET, DET_STRIDE), dtype=dets.dtype, device=dets.device)
    pc = torch.zeros((per,), dtype=counts.dtype, device=counts.device)
    pd[:m] = dets
    pc[:m] = counts
    out_d = [torch.empty_like(pd) for _ in range(world)]
    out_c = [torch.empty_like(pc) for _ in range(world)]
    dist.all_gather(out_d, pd, group=group)
    dist.all_gather(out_c, pc, group=group)
    if rank != dst:
        return None, None
    return torch.cat(out_d)[:n_images], torch.cat(out_c)[:n_images]


def reduce_max_a(local_max, group=None):
    """Calibration (stage_4 / stage_5): elementwise max over ranks of the per-tap activation maxima.
    local_max: {tap name: float}; every rank must hold the same keys.  Returns the reduced dict on every rank."""
    keys = sorted(local_max)
    t = torch.tensor([float(local_max[k]) for k in keys], dtype=torch.float32)
    if dist.get_backend(group) == 'nccl':
        t = t.cuda()
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    vals = t.cpu().tolist()
    return {k: v for k, v in zip(keys, vals)}


def bind_to_gpu_numa_node(device_index):
    """Pin this process (one rank per GPU) to the CPUs of the NUMA node its GPU hangs off, BEFORE it allocates pinned host
    buffers: first-touch then places them in the local DRAM, so the H2D copies of the ranks do not cross the socket
    interconnect.  Returns the cpu list it bound to, or None when the topology is not exposed (nothing changes then)."""
    import os
    try:
        props = torch.cuda.get_device_properties(device_index)
        bdf = f'{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0'
        with open(f'/sys/bus/pci/devices/{bdf}/local_cpulist') as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(','):
            if '-' in part:
                lo, hi = part.split('-')
                cpus.update(range(int(lo), int(hi) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except (OSError, AttributeError, ValueError, RuntimeError, AssertionError):
        return None

