"""Per-launch table from an `ncu --csv --page raw` log of the conv kernels of one pass: time, DRAM, L2 bytes, pipe utilisation,
joined with the plan's layer order and algorithmic bytes.   python tools/ncu_layers.py gpurun_out/conv_tma_l2_s4.csv [batch]"""
import csv
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from alpha_yolo_quant_b200 import loaders, plan  # noqa: E402


def main(path, batch=256):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    K, sd, sc, ma = loaders.load_workload_npz(os.path.join(REPO, 'tests', 'golden', 'workload_k8.npz'))
    p = plan.compile_plan(sd, sc, ma, K)
    names = [n for n in p.op_names if n in p.info['layers'] and 'alg_in_bytes' in p.info['layers'][n]]
    col = lambda r, k: float(r[hdr.index(k)].replace(',', ''))
    tot = dict(t=0, l2=0, alg=0, dr=0, dw=0)
    print(f'{"layer":20s} {"us":>7s} {"L2 MB":>8s} {"alg MB":>7s} {"L2/alg":>6s} {"DRAMr":>6s} {"DRAMw":>6s} {"L2 TB/s":>7s} {"tens%":>5s} {"issue%":>6s}')
    for nm, r in zip(names, data):
        L = p.info['layers'][nm]
        t = col(r, 'gpu__time_duration.sum')
        t = t / 1e3 if t > 5000 else t          # ns -> us
        l2 = col(r, 'lts__t_bytes.sum'); u = r and rows[1][hdr.index('lts__t_bytes.sum')]
        l2 = l2 * {'Gbyte': 1e3, 'Mbyte': 1.0, 'Kbyte': 1e-3, 'byte': 1e-6}.get(u, 1.0)
        conv = lambda k: col(r, k) * {'Gbyte': 1e3, 'Mbyte': 1.0, 'Kbyte': 1e-3, 'byte': 1e-6}.get(rows[1][hdr.index(k)], 1.0)
        dr, dw = conv('dram__bytes_read.sum'), conv('dram__bytes_write.sum')
        alg = (L['alg_in_bytes'] + L['alg_out_bytes']) * batch / 1e6
        print(f'{nm:20s} {t:7.1f} {l2:8.0f} {alg:7.0f} {l2 / alg:6.2f} {dr:6.0f} {dw:6.0f} {l2 / t / 1e0 * 1e-6 * 1e6 / 1e6:7.2f} '
              f'{col(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"):5.1f} {col(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"):6.1f}')
        tot['t'] += t; tot['l2'] += l2; tot['alg'] += alg; tot['dr'] += dr; tot['dw'] += dw
    print(f'TOTAL {tot["t"]:.0f} us, L2 {tot["l2"]:.0f} MB = {tot["l2"] / tot["alg"]:.2f} x algorithmic ({tot["alg"]:.0f} MB), DRAM {tot["dr"]:.0f} + {tot["dw"]:.0f} MB')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 256)
