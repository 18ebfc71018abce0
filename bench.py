#!/usr/bin/env python
"""bench.py -- images/s of the bit-exact integer YOLOv8n forward + q_NMS (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path (per-image input quantiser -> 63 quantised convs -> SPPF / upsample /
concat -> DFL decode -> q_NMS) over one batch of synthetic 640x640 images per GPU (weak scaling: every rank
gets its own batch of B images, no collective on the data path).  `value` = images all ranks processed /
max-over-ranks device time, inputs resident in HBM.  `e2e` = the same through ayq_forward_host with pinned
HOST buffers (H2D of the fp32 images and D2H of the detections inside the timed region).

--impl reference times the CPU restatement of the reference (oracle/yolo_int.py, kind "port"; the reference
itself is pure Python under /root/reference, which does not exist on the GPU box) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = 'images/sec @640x640 bit-exact int YOLOv8n+q_NMS'
OPS_IMG = 2 * 4371456000
BYTES_IMG = 39993600                # fp32 image read + sum conv int8 inputs + outputs
WORKLOAD = ('YOLOv8n full_quant + Detect head + q_NMS (stage_8_torch_full_quant path), K=8, 640x640, batch 256 per GPU '
            '(BASELINE configs[2]), random-init weights through the reference stage_2-7 pipeline')
CONV_ALG_BYTES_IMG = 32211200       # sum over the 62 tcgen05 convs of input read + output written at 1 B / element


def synth_batch_u8(n, seed0=0):
    """n synthetic uint8 CHW images.  A few distinct generator images (oracle/synth.py families) are tiled and
    rolled so every image differs while generation stays fast."""
    from oracle import synth
    base = [synth.synth_image_u8(seed0 + s) for s in range(min(n, 8))]
    out = np.empty((n, 3, 640, 640), np.uint8)
    for i in range(n):
        out[i] = np.roll(base[i % len(base)], shift=(7 * (i // len(base)), 13 * (i // len(base))), axis=(1, 2))
    return out


def load_peaks():
    p = os.path.join(REPO, 'MEASURED_PEAKS.json')
    out = dict(hbm=6650.0, bf16=1590.0, bf16_sus=1400.0, src='fallback')
    if os.path.exists(p):
        d = json.load(open(p))
        out = dict(hbm=d['hbm_gbs'], bf16=d['bf16_tflops'], bf16_sus=d.get('bf16_tflops_sustained', d['bf16_tflops']), src='measured')
    # dense int8 tcgen05 peak of this pool's B200 measured by tools/ubench/ubench.cu (kind::i8, M=128, N=256, K=32 from resident smem)
    q = os.path.join(REPO, 'profiles', 'int8_peak_r2.json')
    if os.path.exists(q):
        j = json.load(open(q))
        out.update(int8=j['int8_tops_dense'], int8_src=j['source'])
    else:
        out.update(int8=2 * out['bf16_sus'], int8_src='2 x sustained bf16 (no int8 probe committed)')
    return out


class ClockSampler:
    FIELDS = 'clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(['nvidia-smi', f'--id={index}', f'--query-gpu={self.FIELDS}', '--format=csv,noheader,nounits',
                                       '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return None
        time.sleep(0.15)
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:
            self.p.kill()
            return None
        sm, mx, reasons = [], 0, set()
        for line in out.splitlines():
            parts = [s.strip() for s in line.split(',')]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx = max(mx, float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), parts[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        if not sm:
            return None
        return {'sm_mhz': float(np.median(sm)), 'sm_max_mhz': mx, 'reasons': sorted(reasons), 'samples': len(sm)}


_ORACLE = None


def _oracle_init():
    """Worker initialiser: one oracle per process, single-threaded maths (the processes are the parallelism)."""
    global _ORACLE
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    torch.set_num_threads(1)
    from oracle import yolo_int as Y
    _ORACLE = Y.OracleYolov8(Y.Workload(os.path.join(REPO, 'tests', 'golden', 'workload_k8.npz')))


def _oracle_run(seed):
    from oracle import synth
    x = synth.to_input_array([synth.synth_image_u8(seed)])
    res = _ORACLE.forward(x)
    return 0 if res[0][0] is None else int(res[0][0].shape[0])


class CpuReference:
    """The reference's CPU path (numpy restatement oracle/yolo_int.py, batch 1 per call like the reference driver loop,
    stage_8_torch.py:1004-1013) on all host cores: one worker process per core, each with its own model."""

    def __init__(self, workers=None):
        import multiprocessing as mp
        from concurrent.futures import ProcessPoolExecutor
        self.workers = workers or min(os.cpu_count() or 1, 32)     # bounded: each worker holds ~0.5 GB of im2col scratch
        saved = {k: os.environ.get(k) for k in ('OMP_NUM_THREADS', 'MKL_NUM_THREADS', 'OPENBLAS_NUM_THREADS')}
        for k in saved:                                            # the workers are the parallelism: one BLAS thread each
            os.environ[k] = '1'                                    # (must be in the environment BEFORE the children import numpy)
        self.pool = ProcessPoolExecutor(max_workers=self.workers, mp_context=mp.get_context('spawn'), initializer=_oracle_init)
        list(self.pool.map(_oracle_run, range(self.workers)))          # start every worker, build its LUTs, warm caches
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v

    def rate(self, n_images, seed0=100):
        t0 = time.perf_counter()
        dets = list(self.pool.map(_oracle_run, range(seed0, seed0 + n_images)))
        dt = time.perf_counter() - t0
        return n_images / dt, dt, sum(dets)

    def close(self):
        self.pool.shutdown()


def run_reference(args, rank):
    """--impl reference: the reference's own CPU implementation of the path on the box's host cores.  The reference is
    pure Python under /root/reference, which does not exist on the GPU box, so this is the pinned port (kind "port")."""
    if rank != 0:
        return
    ref = CpuReference()
    per_step = 2 * ref.workers                       # bounded sample: two images per core per step
    for _ in range(min(args.warmup, 1)):
        ref.rate(ref.workers)
    t_total, n_total = 0.0, 0
    for _ in range(args.steps):
        _, dt, _ = ref.rate(per_step)
        t_total += dt; n_total += per_step
    ref.close()
    value = n_total / t_total
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'images/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1000.0 * t_total / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'int8 weights/activations, int32 accumulate, fp32-rounded requant products', 'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'sample': f'{per_step} images per step, batch 1 per call, one worker process per host core'},
        'cpu_baseline': {'value': value, 'unit': 'images/s', 'cores': ref.workers, 'kind': 'port',
                         'sample': f'{n_total} synthetic images, numpy oracle of stage_8_torch_full_quant (oracle/yolo_int.py), '
                                   f'{ref.workers} worker processes x batch 1'},
        'e2e': {'value': value, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batch', type=int, default=256, help='images per GPU per step')
    ap.add_argument('--max-batch', type=int, default=512, help='images per internal pass of the engine')
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--conv', default='tma', choices=['tma'], help='the product library has one convolution kernel family')
    ap.add_argument('--cpu-images', type=int, default=8, help='bounded CPU-baseline sample (images)')
    ap.add_argument('--k', type=int, default=8, choices=[8, 6, 4], help='bit width of weights / activations (BASELINE configs[4]: 6 / 4)')
    ap.add_argument('--sustain', type=float, default=0.0, help='additionally run the device-resident loop for this many seconds (clock record)')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--ops-json', default=None, help='write the per-op time table here')
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'ours':
        args.warmup = 3
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if args.impl == 'reference':
        run_reference(args, rank)
        return

    from alpha_yolo_quant_b200 import engine, loaders, plan
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the product path has no CPU fallback')
    torch.cuda.set_device(local)
    if world > 1:                                                  # host buffers of a rank live next to its GPU (NUMA)
        from alpha_yolo_quant_b200 import dataparallel as _dp
        _dp.bind_to_gpu_numa_node(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    K, sd, sc, ma = loaders.load_workload_npz(os.path.join(REPO, 'tests', 'golden', f'workload_k{args.k}.npz'))
    p = plan.compile_plan(sd, sc, ma, K)
    # product-level data-parallel entry (alpha_yolo_quant_b200/dataparallel.py): one engine per rank under torchrun
    from alpha_yolo_quant_b200 import dataparallel as dpar
    dpy = dpar.DataParallelYolo(p, devices=[local], max_batch=args.max_batch, group=dist.group.WORLD if dist is not None else None)
    e = dpy.engines[0]
    B = args.batch
    u8 = synth_batch_u8(B, seed0=17 * rank)
    host_u8 = torch.from_numpy(u8).pin_memory()
    host_f32 = (torch.from_numpy(u8).float() / 255.0).pin_memory()
    x = host_f32.cuda(non_blocking=True)
    dets = torch.empty((B, 300, 6), dtype=torch.float32, device='cuda')
    counts = torch.empty((B,), dtype=torch.int32, device='cuda')
    stream = torch.cuda.current_stream()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing (value)
    for _ in range(args.warmup):
        e.forward_into(x, dets, counts)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        e.forward_into(x, dets, counts)
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms], dtype=torch.float64, device='cuda')
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = B * world * args.steps / (ms_max / 1000.0)
    n_det = int(counts.sum().item())
    # ---- the same loop with the images resident as uint8 (the loader's format before ToTensor: ayq_forward_u8); reported beside
    # `value`, which stays on the reference forward()'s own input format (float32 in [0,1])
    x_u8 = torch.from_numpy(u8).cuda()
    for _ in range(args.warmup):
        e.forward_into(x_u8, dets, counts)
    barrier()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record(stream)
    for _ in range(args.steps):
        e.forward_into(x_u8, dets, counts)
    ev3.record(stream)
    barrier()
    t8 = torch.tensor([ev2.elapsed_time(ev3)], dtype=torch.float64, device='cuda')
    if dist is not None:
        dist.all_reduce(t8, op=dist.ReduceOp.MAX)
    value_u8 = {'value': B * world * args.steps / (float(t8.item()) / 1000.0), 'unit': 'images/s', 'ms_per_step': float(t8.item()) / args.steps,
                'input': 'uint8 (B,3,640,640) resident in HBM -> ayq_forward_u8 (ToTensor + input quantiser inside Conv_P1); same detections',
                'same_detections': int(counts.sum().item()) == n_det}
    del x_u8
    # ---- end to end through the host-buffer C-ABI call
    e2e = None
    e2e_f32 = None
    if not args.no_e2e:
        # The call a user makes: DataParallelYolo / Engine.forward_host_async with pinned HOST buffers, one call per step, all K
        # steps queued behind one another (each with its own result buffers) and ONE wait at the end, as a validation driver that
        # streams batches would do (stage_8_torch.py:1004-1013 re-hosted).  Every step's H2D of the images and D2H of its detections
        # are inside the timed region; consecutive steps overlap (upload of step i+1 under the kernels of step i), so only the first
        # upload and the last pass of the whole run are exposed.  `sync_value` is the same with a blocking call per step.
        NB = min(args.steps, 4)
        dets_h = [torch.empty((B, 300, 6), dtype=torch.float32).pin_memory() for _ in range(NB)]
        counts_h = [torch.empty((B,), dtype=torch.int32).pin_memory() for _ in range(NB)]
        res, res_sync = {}, {}
        for name, src in (('f32', host_f32), ('u8', host_u8)):
            for _ in range(2):
                e.forward_host(src, dets_h[0], counts_h[0])
            barrier()
            t0 = time.perf_counter()
            for i in range(args.steps):
                e.forward_host_async(src, dets_h[i % NB], counts_h[i % NB])
            e.wait()
            dt = time.perf_counter() - t0
            tt = torch.tensor([dt], dtype=torch.float64, device='cuda')
            if dist is not None:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            res[name] = B * world * args.steps / float(tt.item())
            for j in range(NB):
                assert int(counts_h[j].sum().item()) == n_det
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                dpy.forward_shard(src, dets_h[0], counts_h[0])     # blocking per step: returns with the results on the host
            dt = time.perf_counter() - t0
            tt = torch.tensor([dt], dtype=torch.float64, device='cuda')
            if dist is not None:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            res_sync[name] = B * world * args.steps / float(tt.item())
            assert int(counts_h[0].sum().item()) == n_det
        d2h = int(dets_h[0].numel() * 4 + counts_h[0].numel() * 4)
        # Headline: uint8 host images, the format the reference's validation loader holds before ToTensor
        # (stage_8_torch.py:985-990): ToTensor + forward + q_NMS on the GPU.  `e2e_f32` is the reference's own forward() input
        # format (fp32 (N,3,640,640), stage_8_torch_full_quant.py:704-710; 4x the PCIe bytes) measured the same way.
        e2e = {'value': res['u8'], 'unit': 'images/s', 'h2d_bytes_per_step': int(host_u8.numel()), 'd2h_bytes_per_step': d2h,
               'input': 'uint8 (B,3,640,640) pinned host -> ayq_forward_host_async x steps + ayq_wait (ToTensor on the GPU)',
               'sync_value': res_sync['u8'], 'f32_input_value': res['f32'], 'f32_h2d_bytes_per_step': int(host_f32.numel() * 4)}
        e2e_f32 = {'value': res['f32'], 'unit': 'images/s', 'h2d_bytes_per_step': int(host_f32.numel() * 4), 'd2h_bytes_per_step': d2h,
                   'input': 'float32 (B,3,640,640) pinned host (the reference forward() input) -> ayq_forward_host_async x steps + ayq_wait',
                   'sync_value': res_sync['f32']}
        h2d_probe = os.path.join(REPO, 'profiles', 'h2d_probe_r2.json')
        if os.path.exists(h2d_probe):                              # raw concurrent cudaMemcpyAsync H2D rate of this pool's box at N GPUs
            hp = json.load(open(h2d_probe)).get(str(world))
            if hp:
                e2e['h2d_gbs'] = res['u8'] * host_u8[0].numel() / 1e9
                e2e['h2d_probe_gbs'] = hp
                e2e['frac_of_h2d_probe'] = e2e['h2d_gbs'] / hp
                e2e_f32['h2d_gbs'] = res['f32'] * host_f32[0].numel() * 4 / 1e9
                e2e_f32['frac_of_h2d_probe'] = e2e_f32['h2d_gbs'] / hp

    # ---- every rank's detections equal the single-GPU result (rank 0 recomputes each rank's batch on its own GPU; outside timing)
    parity = None
    if dist is not None:
        gd, gc = dpy.gather(dets.cpu(), counts.cpu(), B * world)
        if rank == 0:
            bad = 0
            for r in range(world):
                xr = (torch.from_numpy(synth_batch_u8(B, seed0=17 * r)).float() / 255.0).cuda()
                dr, cr = e.forward(xr)
                dr, cr = dr.cpu(), cr.cpu()
                if not torch.equal(cr, gc[r * B:(r + 1) * B]):
                    bad += 1
                    continue
                for i in range(B):
                    k = int(cr[i])
                    if not torch.equal(dr[i, :k], gd[r * B + i, :k]):
                        bad += 1
                        break
            parity = {'ranks_checked': world, 'ranks_differing': bad, 'images': B * world,
                      'check': 'rank 0 recomputed every rank\'s batch on GPU 0: counts and all detection rows bit-identical'}
            assert bad == 0, parity

    # ---- sustained run (clock record): the device-resident loop for >= args.sustain seconds
    sustained = None
    if args.sustain > 0 and rank == 0 and world == 1:
        sampler2 = ClockSampler(local)
        n_iter = max(int(args.sustain / (ms_max / args.steps / 1000.0)), 1)
        evs0, evs1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        evs0.record(stream)
        for _ in range(n_iter):
            e.forward_into(x, dets, counts)
        evs1.record(stream)
        torch.cuda.synchronize()
        ms_s = evs0.elapsed_time(evs1)
        sustained = {'seconds': ms_s / 1000.0, 'steps': n_iter, 'value': B * n_iter / (ms_s / 1000.0), 'unit': 'images/s', 'clocks': sampler2.stop()}

    # ---- per-op device times (CUDA events around every kernel, outside the timed region) -> roofline of the dominant kernel
    roofline, top = None, None
    if rank == 0:
        e.set_profiling(True)
        for _ in range(3):
            e.forward_into(x, dets, counts)
        torch.cuda.synchronize()
        op_ms, op_calls = e.op_times()
        e.set_profiling(False)
        peaks = load_peaks()
        rows = []
        names = ['absmax(quant_matrix)'] + list(p.op_names)
        passes_per_step = (B + args.max_batch - 1) // args.max_batch
        for i in range(len(op_ms)):
            if op_calls[i] == 0:
                continue
            avg_ms = op_ms[i] / op_calls[i]
            nm = names[i] or f'op{i - 1}'
            meta = p.info['layers'].get(nm)
            imgs = min(B, args.max_batch)
            row = {'op': nm, 'avg_ms': float(avg_ms), 'share': float(op_ms[i] / op_ms.sum())}
            if meta:
                out_b, in_b = meta['out_bytes'], meta['in_bytes']
                alg = meta.get('alg_in_bytes', in_b) + meta.get('alg_out_bytes', out_b)    # independent of layout / fusion
                row.update(macs_per_img=meta['macs'], bytes_per_img=alg, stored_bytes_per_img=in_b + out_b,
                           tops=2e-9 * meta['macs'] * imgs / avg_ms, gbs=1e-6 * alg * imgs / avg_ms)
            rows.append(row)
        rows.sort(key=lambda r: -r['avg_ms'])
        top = rows[0]
        # Dominant kernel = the convolution kernel (one kernel, 62 launches per pass, ~85 % of the pass).  SURVEY.md 8(d):
        # the network is HBM-bound at 1 B/activation (219 OP/B against a ~500 OP/B ridge), so the roofline is algorithmic
        # bytes (every conv input plane read once + every output plane written once) over the measured launch time.
        conv_rows = [r for r in rows if 'macs_per_img' in r and r['op'] != 'Conv_P1']
        if conv_rows:
            imgs = min(B, args.max_batch)
            t_ms = sum(r['avg_ms'] for r in conv_rows)
            byt = sum(r['bytes_per_img'] for r in conv_rows) * imgs            # = 32,211,200 B per image x images
            byt_stored = sum(r['stored_bytes_per_img'] for r in conv_rows) * imgs   # what this plan really moves (fused upsample copies, 2 B class logits)
            if args.k == 8:
                assert byt == CONV_ALG_BYTES_IMG * imgs, (byt, CONV_ALG_BYTES_IMG * imgs)
            mac = sum(r['macs_per_img'] for r in conv_rows) * imgs
            gbs = 1e-6 * byt / t_ms
            traffic, traffic_src = None, None
            tp_path = os.path.join(REPO, 'profiles', 'traffic_r2.json')
            if os.path.exists(tp_path):     # DRAM bytes of all conv launches of one pass, ncu with caches left alone (--cache-control none)
                tj = json.load(open(tp_path))
                traffic = tj['dram_bytes_per_launch'] * imgs / tj['images_per_pass']
                traffic_src = tj.get('source')
            roofline = {'bound': 'hbm', 'achieved': gbs, 'peak': peaks['hbm'], 'unit': 'GB/s', 'frac': gbs / peaks['hbm'],
                        'traffic': traffic, 'traffic_source': traffic_src, 'kernel': 'conv_tma_kernel', 'launches_per_pass': len(conv_rows),
                        'avg_launch_us': 1e3 * t_ms / len(conv_rows), 'algorithmic_bytes_per_launch': byt / len(conv_rows),
                        'achieved_with_stored_bytes': 1e-6 * byt_stored / t_ms,
                        'share_of_pass': float(sum(r['share'] for r in conv_rows)), 'peak_source': peaks['src'],
                        'tensor': {'achieved': 2e-9 * mac / t_ms, 'peak': peaks['int8'], 'unit': 'TOP/s dense int8', 'peak_source': peaks['int8_src'],
                                   'frac': 2e-9 * mac / t_ms / peaks['int8']},
                        'note': 'aggregate over all launches of the kernel in one pass: sum of SURVEY 8(d) algorithmic bytes / sum of event-timed durations'}
        if args.ops_json:
            json.dump({'rows': rows, 'batch_per_pass': min(B, args.max_batch), 'conv': args.conv}, open(args.ops_json, 'w'), indent=1)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        ref = CpuReference()
        # bounded sample of the same workload: chunks of two images per core until >= ~12 s of CPU work (at most 60 s)
        chunk = max(args.cpu_images, 2 * ref.workers)
        n_cpu, dt = 0, 0.0
        while dt < 12.0 and n_cpu < 4096:
            _, d, _ = ref.rate(chunk, seed0=100 + n_cpu)
            n_cpu += chunk; dt += d
        r = n_cpu / dt
        ref.close()
        cpu = {'value': r, 'unit': 'images/s', 'cores': ref.workers, 'kind': 'port',
               'sample': f'{n_cpu} synthetic images in {dt:.1f} s, numpy oracle of stage_8_torch_full_quant (oracle/yolo_int.py), '
                         f'{ref.workers} worker processes x batch 1'}

    if rank == 0:
        passes = (B + args.max_batch - 1) // args.max_batch
        peaks = load_peaks()
        line = {
            'metric': METRIC, 'value': value, 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_max / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'int8 weights/activations, int32 accumulate, fp32-rounded requant products', 'data': 'synthetic',
            'config': {'workload': (WORKLOAD if B == 256 else WORKLOAD.replace('batch 256', f'batch {B}')).replace('K=8', f'K={args.k}'),
                       'global_batch': B * world, 'images_per_pass': min(B, args.max_batch), 'conv_kernel': args.conv,
                       'l2': f'inputs larger than L2 ({B * 4915200 / 1e6:.0f} MB fp32 images per step, activations {e.workspace_bytes / 1e6:.0f} MB workspace)',
                       'detections_per_step': n_det},
            'value_u8_resident': value_u8, 'e2e': e2e, 'e2e_f32': e2e_f32, 'gpu_launches': int(e.launches_per_pass * passes * args.steps),
            'clocks': clocks, 'roofline': roofline, 'cpu_baseline': cpu, 'multi_gpu_parity': parity, 'sustained': sustained,
            'whole_net': {'hbm_frac': value / world * BYTES_IMG / 1e9 / peaks['hbm'], 'int8_tops': value / world * OPS_IMG / 1e12,
                          'tensor_frac': value / world * OPS_IMG / 1e12 / peaks['int8'], 'peaks': peaks['src'], 'int8_peak_source': peaks['int8_src']},
        }
        print(json.dumps(line))
    dpy.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
