"""Turn the raw evidence tools/make_profiles.sh leaves in gpurun_out/ into the tracked summaries under profiles/.

  python tools/profiles_from_csv.py r1 [images_per_pass]

Reads  gpurun_out/conv_tma_full_<R>.csv   (ncu --page raw --csv of the 62 conv launches of one pass, few metrics)
       gpurun_out/conv_tma_set_full_<R>.ncu-rep (ncu --set full of three representative launches, optional)
       gpurun_out/ops_<R>.json            (event-timed per-op table of the plain bench run: layer names / algorithmic bytes)
Writes profiles/conv_tma_ncu_full_<R>.md, profiles/traffic_<R>.json and copies launches / ops / bench / role profile.
"""
import csv
import json
import os
import shutil
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(REPO, 'gpurun_out')
P = os.path.join(REPO, 'profiles')


def read_raw_csv(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    return [dict(zip(hdr, r)) for r in rows[2:]]


def main():
    R = sys.argv[1] if len(sys.argv) > 1 else 'r1'
    imgs = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    rows = read_raw_csv(os.path.join(G, f'conv_tma_full_{R}.csv'))
    ops = json.load(open(os.path.join(G, f'ops_{R}.json')))
    conv_ops = [r for r in ops['rows'] if 'bytes_per_img' in r and r['op'] != 'Conv_P1']
    # launch order of the convs in a pass = plan order; ops json is sorted by time, so re-sort by the plan order kept in 'idx'
    sys.path.insert(0, REPO)
    from alpha_yolo_quant_b200.plan import LAYER_INDEX            # launch order of the convs in a pass = plan order
    conv_ops.sort(key=lambda r: LAYER_INDEX[r['op']])
    names = [r['op'] for r in conv_ops] if len(conv_ops) == len(rows) else [f'conv {i}' for i in range(len(rows))]
    alg = sum(r['bytes_per_img'] for r in conv_ops) * imgs
    f = lambda r, k: float(r[k])
    tot_us = sum(f(r, 'gpu__time_duration.sum') for r in rows) / 1e3
    rd = sum(f(r, 'dram__bytes_read.sum') for r in rows)
    wr = sum(f(r, 'dram__bytes_write.sum') for r in rows)
    n = len(rows)
    out = [f'# ncu per-launch metrics: all {n} conv_tma_kernel launches of one {imgs}-image pass ({R})', '',
           f'Command: `AYQ_NO_GRAPH=1 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,'
           f'sm__pipe_tensor_cycles_active...,smsp__issue_active...,launch__* --clock-control none -k regex:conv_tma -s {n} -c {n} '
           f'python tools/one_pass.py --batch {imgs} --passes 2 --conv tma` (tools/make_profiles.sh).',
           '(per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes; the bench line uses CUDA events).', '',
           f'Totals over the {n} launches: duration {tot_us:.0f} us, DRAM read {rd / 1e6:.0f} MB, DRAM write {wr / 1e6:.0f} MB -> '
           f'traffic per launch {(rd + wr) / n / 1e6:.1f} MB (algorithmic {alg / n / 1e6:.1f} MB per launch), '
           f'{(rd + wr) / tot_us / 1e3:.0f} GB/s DRAM under ncu.', '',
           '| layer | kernel | us | DRAM rd MB | DRAM wr MB | L2 MB | tensor pipe % | issue active % | regs | block | grid | smem KB |',
           '|---|---|---|---|---|---|---|---|---|---|---|---|']
    for nm, r in zip(names, rows):
        kn = r['Kernel Name'].split('(')[0]
        out.append(f"| {nm} | {kn} | {f(r, 'gpu__time_duration.sum') / 1e3:.1f} | {f(r, 'dram__bytes_read.sum') / 1e6:.1f} | "
                   f"{f(r, 'dram__bytes_write.sum') / 1e6:.1f} | {f(r, 'lts__t_bytes.sum') / 1e6:.0f} | "
                   f"{f(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | "
                   f"{f(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | {r['launch__registers_per_thread']} | "
                   f"{r['launch__block_size']} | {r['launch__grid_size']} | {f(r, 'launch__shared_mem_per_block_dynamic') / 1024:.0f} |")
    rep = os.path.join(G, f'conv_tma_set_full_{R}.ncu-rep')
    if os.path.exists(rep):
        out += ['', f'## ncu --set full --import-source on: three representative launches (`gpurun_out/conv_tma_set_full_{R}.ncu-rep`)', '',
                '```']
        txt = subprocess.run([sys.executable, os.path.join(REPO, 'tools', 'ncu_summary.py'), rep], capture_output=True, text=True).stdout
        out += txt.splitlines()[:150]
        out.append('```')
    open(os.path.join(P, f'conv_tma_ncu_full_{R}.md'), 'w').write('\n'.join(out) + '\n')
    json.dump({'kernel': 'conv_tma_kernel', 'launches': n, 'images_per_pass': imgs, 'dram_bytes_per_launch': (rd + wr) / n,
               'algorithmic_bytes_per_launch': alg / n, 'source': f'profiles/conv_tma_ncu_full_{R}.md'},
              open(os.path.join(P, f'traffic_{R}.json'), 'w'), indent=1)
    for src, dst in ((f'launches_{R}.csv', f'launches_{R}.csv'), (f'ops_{R}.json', f'ops_{R}.json'), (f'bench_{R}.json', f'bench_{R}.json'),
                     (f'role_profile_{R}.txt', f'role_profile_{R}.txt')):
        if os.path.exists(os.path.join(G, src)):
            shutil.copy(os.path.join(G, src), os.path.join(P, dst))
    print(open(os.path.join(P, f'traffic_{R}.json')).read())


if __name__ == '__main__':
    main()
