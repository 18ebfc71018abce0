"""N > 1 host logic on CPU: world_size-2 gloo process group (no GPU needed).  Sharding, ordered gather of the per-image
detections, and the calibration max-reduction."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from alpha_yolo_quant_b200 import dataparallel as dp


def test_shard_range_covers_every_image_once():
    for n in (0, 1, 2, 7, 8, 9, 255, 256, 4096):
        for g in (1, 2, 3, 4, 8):
            seen = []
            for r in range(g):
                lo, hi = dp.shard_range(n, g, r)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi))
            assert seen == list(range(n)), (n, g)
            assert max(dp.shard_range(n, g, r)[1] - dp.shard_range(n, g, r)[0] for r in range(g)) == (n + g - 1) // g


def _fake_result(i):
    """Deterministic stand-in for the detections of image i (count depends on i, rows carry i)."""
    k = (7 * i) % 11
    d = torch.zeros((300, 6))
    d[:k] = torch.arange(k, dtype=torch.float32).reshape(-1, 1) + 1000.0 * i
    return d, k


def _worker(rank, world, port, n_images, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    lo, hi = dp.shard_range(n_images, world, rank)
    dets = torch.stack([_fake_result(i)[0] for i in range(lo, hi)]) if hi > lo else torch.zeros((0, 300, 6))
    counts = torch.tensor([_fake_result(i)[1] for i in range(lo, hi)], dtype=torch.int32)
    D, C = dp.gather_detections(dets, counts, n_images)
    red = dp.reduce_max_a({'conv_p2': 1.0 + rank, 'start': 1.0, 'conv8': 5.0 - rank})
    ok = True
    if rank == 0:
        ok = D.shape == (n_images, 300, 6) and C.tolist() == [_fake_result(i)[1] for i in range(n_images)]
        for i in range(n_images):
            ok = ok and torch.equal(D[i], _fake_result(i)[0])
    else:
        ok = D is None and C is None
    ok = ok and red == {'conv8': 5.0, 'conv_p2': float(world), 'start': 1.0}
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_gather_and_reduce_world_size_2():
    ctx = mp.get_context('spawn')
    for n_images in (5, 8, 1):                        # ragged, even, and a rank with an empty shard
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, 2, port, n_images, q)) for r in range(2)]
        for p in procs:
            p.start()
        res = sorted(q.get(timeout=120) for _ in procs)
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        assert res == [(0, True), (1, True)], (n_images, res)


def test_calibration_text_round_trip():
    """max_a_all.txt / max_a.txt formats (stage_4.py:1007-1011, stage_5.py:11-33, stage_5_common_func.py:11-26)."""
    from alpha_yolo_quant_b200 import calibration as cal
    from alpha_yolo_quant_b200.plan import parse_max_a
    m = {'start': [torch.tensor(1.0), torch.tensor(1.0)],
         'conv_p2': [torch.tensor(1.42713), torch.tensor(0.5)],
         'conv8': [torch.tensor(12.00004), torch.tensor(33.25)]}
    txt = cal.format_max_a_all(m)
    assert txt.splitlines()[1] == 'conv_p2: [tensor(1.4271), tensor(0.5000)]'
    back = cal.parse_max_a_all(txt)
    assert back == {'start': [1.0, 1.0], 'conv_p2': [1.4271, 0.5], 'conv8': [12.0, 33.25]}
    out = cal.format_max_a(back)
    assert out == 'start: 1.0\nconv_p2: 1.4271\nconv8: 33.25\n'
    assert parse_max_a(out) == {'start': 1.0, 'conv_p2': 1.4271, 'conv8': 33.25}


def test_numa_binding_is_a_no_op_without_topology():
    """no GPU / no sysfs topology -> returns None and leaves the affinity untouched (the bench calls it on every rank)"""
    import os
    from alpha_yolo_quant_b200 import dataparallel as dp
    before = os.sched_getaffinity(0)
    r = dp.bind_to_gpu_numa_node(0)
    assert r is None or set(r) <= before
    if r is None:
        assert os.sched_getaffinity(0) == before


class _StubEngine:
    """CPU stand-in with the Engine methods DataParallelYolo uses; 'detections' encode the image's first pixel and the device."""
    log = []

    def __init__(self, plan, device, max_batch):
        self.device, self.queue = device, []

    def forward_host_async(self, img, dets, counts):
        assert img.is_contiguous() and dets.is_contiguous()
        self.queue.append((img, dets, counts))
        _StubEngine.log.append(('enqueue', self.device, img.shape[0]))

    def wait(self):
        _StubEngine.log.append(('wait', self.device))
        for img, dets, counts in self.queue:
            for i in range(img.shape[0]):
                k = int(img[i, 0, 0, 0]) % 5
                counts[i] = k
                dets[i].zero_()
                dets[i, :k, 0] = float(img[i, 0, 0, 0])
                dets[i, :k, 5] = self.device
        self.queue = []

    def close(self):
        pass


def test_data_parallel_single_process_sharding_and_order():
    """One process, G 'GPUs': contiguous ceil(N / G) shards, every shard queued before the first wait (so the devices overlap),
    results in image order in one pair of arrays; ragged and empty shards."""
    for n, g in ((10, 4), (8, 8), (3, 8), (257, 2), (1, 1)):
        _StubEngine.log = []
        dpy = dp.DataParallelYolo(plan=None, devices=list(range(g)), engine_factory=_StubEngine)
        img = torch.zeros((n, 3, 2, 2), dtype=torch.uint8)
        img[:, 0, 0, 0] = torch.arange(n, dtype=torch.uint8)
        dets, counts = dpy.forward_host(img)
        per = (n + g - 1) // g
        for i in range(n):
            k = (i % 256) % 5
            assert int(counts[i]) == k
            assert dets[i, :k, 0].tolist() == [float(i % 256)] * k and dets[i, :k, 5].tolist() == [float(i // per)] * k
        kinds = [t[0] for t in _StubEngine.log]
        assert kinds == sorted(kinds)                              # all 'enqueue' entries precede the first 'wait'
        assert sum(t[2] for t in _StubEngine.log if t[0] == 'enqueue') == n
        dpy.close()


def _dp_worker(rank, world, port, n_images, q):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    dpy = dp.DataParallelYolo(plan=None, devices=[rank], group=dist.group.WORLD, engine_factory=_StubEngine)
    img = torch.zeros((n_images, 3, 2, 2), dtype=torch.uint8)
    img[:, 0, 0, 0] = torch.arange(n_images, dtype=torch.uint8)
    lo, hi = dp.shard_range(n_images, world, rank)
    d, c = dpy.forward_shard(img[lo:hi])
    D, C = dpy.gather(d, c, n_images)
    ok = True
    if rank == 0:
        per = (n_images + world - 1) // world
        for i in range(n_images):
            k = i % 5
            ok = ok and int(C[i]) == k and D[i, :k, 0].tolist() == [float(i)] * k and D[i, :k, 5].tolist() == [float(i // per)] * k
    else:
        ok = D is None
    q.put((rank, bool(ok)))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_one_process_per_gpu_world_size_2():
    """torchrun layout on CPU (gloo): each rank runs its shard, rank 0 gets all N results in image order."""
    ctx = mp.get_context('spawn')
    for n_images in (7, 2, 1):
        q = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, n_images, q)) for r in range(2)]
        for p in procs:
            p.start()
        res = sorted(q.get(timeout=120) for _ in procs)
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0
        assert res == [(0, True), (1, True)], (n_images, res)
