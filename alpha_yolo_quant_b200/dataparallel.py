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



class DataParallelYolo:
    """N host images in, N results out, over every GPU of the box (north_star: "the batch of images is sharded across the 8 GPUs
    of one box ... each GPU loads its own copy of the weights and results are gathered to the host, with no NCCL needed on the
    inner path").  Re-hosts the reference's validation loop body (stage_8_torch.py:1004-1013: img -> model(img) -> boxes,
    classes on the host) for whole batches.  Two ways to drive it:

      * one process, G GPUs (`devices=[0, 1, ...]`): one engine (= one weight copy + workspace) per device, the batch is cut into
        contiguous shards of ceil(N / G) images, every shard is queued with ayq_forward_host_async on its GPU (the calls only
        enqueue, so one host thread keeps all GPUs busy) and the results land directly in ONE pair of host arrays in image order;
      * one process per GPU under torchrun (`group=`): `forward_shard` runs this rank's shard, `gather` (an all_gather of the
        small result arrays, outside the data path) returns all N results in image order on rank 0.

    `engine_factory(plan, device, max_batch)` exists for the CPU tests (a stub engine); the default builds the CUDA engine and
    fails loudly without one.
    """

    def __init__(self, plan, devices=None, max_batch=512, group=None, engine_factory=None):
        if engine_factory is None:
            from . import engine as _eng
            engine_factory = _eng.Engine
        self.group = group
        if group is not None or (devices is None and dist.is_available() and dist.is_initialized()):
            # one rank per GPU: this process owns the GPU torch has selected for it
            dev = torch.cuda.current_device() if devices is None else devices[0]
            self.devices = [dev]
            self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        else:
            if devices is None:
                devices = list(range(torch.cuda.device_count()))
            if not devices:
                raise RuntimeError('DataParallelYolo: no CUDA device (there is no CPU path)')
            self.devices = list(devices)
            self.rank, self.world = 0, 1
        self.engines = [engine_factory(plan, d, max_batch) for d in self.devices]

    def close(self):
        for e in self.engines:
            e.close()
        self.engines = []

    # ---- one process, G GPUs
    def forward_host(self, images, dets=None, counts=None):
        """images: host tensor (N,3,640,640), uint8 (the loader's format before ToTensor, stage_8_torch.py:985-990) or float32 in
        [0,1]; pinned for full speed.  Returns (dets (N,300,6) float32, counts (N) int32) host tensors (pinned) whose row i is
        the reference's model(images[i:i+1]) (counts[i] == 0 <=> (None, None))."""
        n = images.shape[0]
        if dets is None:
            dets = torch.empty((n, MAX_DET, DET_STRIDE), dtype=torch.float32)
            counts = torch.empty((n,), dtype=torch.int32)
            if torch.cuda.is_available():
                dets, counts = dets.pin_memory(), counts.pin_memory()
        g = len(self.engines)
        busy = []
        for r, e in enumerate(self.engines):
            lo, hi = shard_range(n, g, r)
            if hi > lo:
                e.forward_host_async(images[lo:hi], dets[lo:hi], counts[lo:hi])     # contiguous slices of contiguous tensors
                busy.append(e)
        for e in busy:
            e.wait()
        return dets, counts

    # ---- one process per GPU (torchrun): this rank's shard of a global batch of n_images
    def forward_shard(self, shard_images, dets=None, counts=None):
        """Runs THIS rank's contiguous shard (host tensor, shard_range(n_images, world, rank) of the global batch) on its GPU
        through the pipelined host entry; returns pinned host (dets, counts) of the shard."""
        m = shard_images.shape[0]
        if dets is None:
            dets = torch.empty((m, MAX_DET, DET_STRIDE), dtype=torch.float32)
            counts = torch.empty((m,), dtype=torch.int32)
            if torch.cuda.is_available():
                dets, counts = dets.pin_memory(), counts.pin_memory()
        if m:
            self.engines[0].forward_host_async(shard_images, dets, counts)
            self.engines[0].wait()
        return dets, counts

    def gather(self, dets, counts, n_images, dst=0):
        """All ranks' shard results -> (N,300,6), (N) in image order on rank dst (None, None elsewhere)."""
        if self.world == 1:
            return dets, counts
        if dist.get_backend(self.group) == 'nccl':
            dev = torch.device('cuda', self.devices[0])
            d, c = gather_detections(dets.to(dev), counts.to(dev), n_images, self.group, dst)
            return (d.cpu(), c.cpu()) if d is not None else (None, None)
        return gather_detections(dets, counts, n_images, self.group, dst)
