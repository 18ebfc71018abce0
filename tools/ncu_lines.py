"""Attribute executed warp-instructions / stall samples of one profiled kernel to CUDA source lines.
   python tools/ncu_lines.py <rep> <kernel-regex> <launch index> [mangled-name substring] [libayq.so]
Joins `ncu --page source` (per SASS address) with `nvdisasm -g` line info of the cubin inside the .so."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile


def main(rep, kre, idx, fnpat=None, so='alpha_yolo_quant_b200/libayq.so'):
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(so)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
    dis = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    # per function: offset -> (file, line)
    line_of, cur_fn, cur_line = {}, None, None
    for ln in dis.splitlines():
        m = re.match(r'\s*\.section\s+\.text\.(\S+?),', ln)
        if m:
            cur_fn = m.group(1); cur_line = None; continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2))); continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
        if m and cur_fn:
            line_of[(cur_fn, int(m.group(1), 16))] = cur_line
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kre], capture_output=True, text=True).stdout
    blocks, cur = [], None
    for r in csv.reader(out.splitlines()):
        if r and r[0] == 'Kernel Name':
            cur = []; blocks.append((r[1], cur))
        elif cur is not None:
            cur.append(r)
    name, b = blocks[int(idx)]
    h, data = b[0], b[1:]
    iA, iI, iSm = h.index('Address'), h.index('Instructions Executed'), h.index('# Samples')
    base = int(data[0][iA], 16)
    fn = [f for (f, _) in line_of if (fnpat or kre.split('|')[0]) in f]
    fn = fn[0] if fn else None
    ex, sm = collections.Counter(), collections.Counter()
    for r in data:
        key = line_of.get((fn, int(r[iA], 16) - base))
        ex[key] += int(r[iI]); sm[key] += int(r[iSm])
    tot, tots = sum(ex.values()), sum(sm.values())
    print(f'{name[:60]}: {tot} warp instr, {tots} samples')
    srcs = {}
    for key, c in ex.most_common(40):
        txt = ''
        if key:
            for root in ('alpha_yolo_quant_b200/csrc',):
                p = os.path.join(root, key[0])
                if os.path.exists(p):
                    srcs.setdefault(p, open(p).read().splitlines())
                    txt = srcs[p][key[1] - 1].strip()[:90]
        print(f'{100 * c / tot:5.1f}% exec {100 * sm[key] / max(tots, 1):5.1f}% samp  {key}  {txt}')


if __name__ == '__main__':
    main(*sys.argv[1:])
