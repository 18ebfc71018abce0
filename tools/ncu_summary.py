"""Summarise an .ncu-rep: key metrics per launch + top SASS opcodes / stall hot spots (reads with `ncu -i`)."""
import collections
import csv
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed.sum', 'sm__cycles_elapsed.max',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio' ]


def run(args):
    return subprocess.run(['ncu'] + args, capture_output=True, text=True).stdout


def main(path, topn=12):
    rows = list(csv.reader(run(['-i', path, '--page', 'raw', '--csv']).splitlines()))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('====', r[hdr.index('Kernel Name')][:60], 'id', r[0])
        for w in WANT:
            if w in hdr:
                print(f'   {w:70s} {r[hdr.index(w)]} {units[hdr.index(w)]}')
        # stall reasons
        st = [(float(r[i]), h) for i, h in enumerate(hdr) if h.startswith('smsp__average_warps_issue_stalled_') and h.endswith('_per_issue_active.ratio') and r[i]]
        st.sort(reverse=True)
        print('   stalls/issue:', ', '.join(f"{h.split('stalled_')[1].split('_per_issue')[0]}={v:.2f}" for v, h in st[:7]))
    src = list(csv.reader(run(['-i', path, '--page', 'source', '--csv']).splitlines()))
    blocks, cur = [], None
    for r in src:
        if r and r[0] == 'Kernel Name':
            cur = []
            blocks.append((r[1], cur))
        elif cur is not None:
            cur.append(r)
    for name, b in blocks:
        h, data = b[0], b[1:]
        iS, iI, iSm = h.index('Source'), h.index('Instructions Executed'), h.index('# Samples')
        tot = sum(int(r[iI]) for r in data)
        tots = sum(int(r[iSm]) for r in data)
        print('==== source', name[:50], 'warp instr', tot, 'samples', tots)
        op = collections.Counter(); sm = collections.Counter()
        for r in data:
            toks = r[iS].split()
            o = toks[1] if toks[0].startswith('@') else toks[0]
            op[o] += int(r[iI]); sm[o] += int(r[iSm])
        print('   by executed:', ', '.join(f'{o}={100 * c / tot:.1f}%' for o, c in op.most_common(topn)))
        print('   by samples :', ', '.join(f'{o}={100 * c / max(tots, 1):.1f}%' for o, c in sm.most_common(topn)))


if __name__ == '__main__':
    main(sys.argv[1])
