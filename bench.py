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
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d['hbm_gbs'], bf16=d['bf16_tflops'], bf16_sus=d.get('bf16_tflops_sustained', d['bf16_tflops']), src='measured')
    return dict(hbm=6650.0, bf16=1590.0, bf16_sus=1400.0, src='fallback')


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
    ap.add_argument('--max-batch', type=int, default=256, help='images per internal pass of the engine')
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--conv', default='tma', choices=['tma'], help='the product library has one convolution kernel family')
    ap.add_argument('--cpu-images', type=int, default=8, help='bounded CPU-baseline sample (images)')
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
    K, sd, sc, ma = loaders.load_workload_npz(os.path.join(REPO, 'tests', 'golden', 'workload_k8.npz'))
    p = plan.compile_plan(sd, sc, ma, K)
    e = engine.Engine(p, local, args.max_batch)
    e.set_conv_impl(args.conv)
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
    if dist is not None:                              # results of all ranks in image order on rank 0 (outside the timed region)
        from alpha_yolo_quant_b200 import dataparallel as dp
        gd, gc = dp.gather_detections(dets, counts, B * world)
        if rank == 0:
            assert gd.shape[0] == B * world
            n_det_global = int(gc.sum().item())

    # ---- end to end through the host-buffer C-ABI call
    e2e = None
    if not args.no_e2e:
        dets_h = torch.empty((B, 300, 6), dtype=torch.float32).pin_memory()
        counts_h = torch.empty((B,), dtype=torch.int32).pin_memory()
        res = {}
        for name, src in (('f32', host_f32), ('u8', host_u8)):
            for _ in range(2):
                e.forward_host(src, dets_h, counts_h)
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                e.forward_host(src, dets_h, counts_h)       # synchronous: returns with results on the host
            dt = time.perf_counter() - t0
            tt = torch.tensor([dt], dtype=torch.float64, device='cuda')
            if dist is not None:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            res[name] = B * world * args.steps / float(tt.item())
            assert int(counts_h.sum().item()) == n_det
        # Headline: uint8 host images, the format the reference's validation loader holds before ToTensor
        # (stage_8_torch.py:985-990, 1004-1013): H2D of the uint8 batch, ToTensor + forward + q_NMS on the GPU, D2H of the
        # detections, all inside the timed C call.  The fp32-host-input variant (4x the PCIe bytes) is reported beside it.
        e2e = {'value': res['u8'], 'unit': 'images/s', 'h2d_bytes_per_step': int(host_u8.numel()),
               'd2h_bytes_per_step': int(dets_h.numel() * 4 + counts_h.numel() * 4),
               'input': 'uint8 (B,3,640,640) pinned host -> ayq_forward_host_u8 (ToTensor on the GPU)',
               'f32_input_value': res['f32'], 'f32_h2d_bytes_per_step': int(host_f32.numel() * 4)}

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
                row.update(macs_per_img=meta['macs'], bytes_per_img=in_b + out_b,
                           tops=2e-9 * meta['macs'] * imgs / avg_ms, gbs=1e-6 * (in_b + out_b) * imgs / avg_ms)
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
            byt = sum(r['bytes_per_img'] for r in conv_rows) * imgs
            mac = sum(r['macs_per_img'] for r in conv_rows) * imgs
            gbs = 1e-6 * byt / t_ms
            kname = {'tcgen05': 'conv_tc_kernel', 'tma': 'conv_tma_kernel'}.get(args.conv, 'conv_dp4a_kernel')
            traffic = None
            tp_path = os.path.join(REPO, 'profiles', 'traffic_r1.json')
            if args.conv == 'tma' and os.path.exists(tp_path):          # DRAM bytes per launch from the committed ncu --set full capture
                tj = json.load(open(tp_path))
                traffic = tj['dram_bytes_per_launch'] * imgs / tj['images_per_pass']
            roofline = {'bound': 'hbm', 'achieved': gbs, 'peak': peaks['hbm'], 'unit': 'GB/s', 'frac': gbs / peaks['hbm'],
                        'traffic': traffic, 'kernel': kname, 'launches_per_pass': len(conv_rows),
                        'avg_launch_us': 1e3 * t_ms / len(conv_rows), 'algorithmic_bytes_per_launch': byt / len(conv_rows),
                        'share_of_pass': float(sum(r['share'] for r in conv_rows)), 'peak_source': peaks['src'],
                        'tensor': {'achieved': 2e-9 * mac / t_ms, 'peak': 2 * peaks['bf16_sus'], 'unit': 'TOP/s int8 (peak = 2 x sustained bf16)',
                                   'frac': 2e-9 * mac / t_ms / (2 * peaks['bf16_sus'])},
                        'note': 'aggregate over all launches of the kernel in one pass: sum of algorithmic bytes / sum of event-timed durations'}
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
            'config': {'workload': WORKLOAD if B == 256 else WORKLOAD.replace('batch 256', f'batch {B}'),
                       'global_batch': B * world, 'images_per_pass': min(B, args.max_batch), 'conv_kernel': args.conv,
                       'l2': f'inputs larger than L2 ({B * 4915200 / 1e6:.0f} MB fp32 images per step, activations {e.workspace_bytes / 1e6:.0f} MB workspace)',
                       'detections_per_step': n_det if dist is None else n_det_global},
            'e2e': e2e, 'gpu_launches': int(e.launches_per_pass * passes * args.steps),
            'clocks': clocks, 'roofline': roofline, 'cpu_baseline': cpu,
            'whole_net': {'hbm_frac': value / world * BYTES_IMG / 1e9 / peaks['hbm'], 'int8_tops': value / world * OPS_IMG / 1e12,
                          'tensor_frac_vs_2x_bf16': value / world * OPS_IMG / 1e12 / (2 * peaks['bf16_sus']), 'peaks': peaks['src']},
        }
        print(json.dumps(line))
    e.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
