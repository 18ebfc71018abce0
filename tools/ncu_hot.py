"""Print the SASS instructions of one profiled launch that collect the most stall samples, with neighbours' source lines.
   python tools/ncu_hot.py <rep> <kernel-regex> <block index> <mangled-name substring> [top]"""
import csv, os, re, subprocess, sys, tempfile

def main(rep, kre, idx, fnpat, top=25, so='alpha_yolo_quant_b200/libayq.so'):
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(so)], cwd=tmp, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith('.cubin')][0]
    dis = subprocess.run(['nvdisasm', '-g', '-c', os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    line_of, cur_fn, cur_line = {}, None, None
    for ln in dis.splitlines():
        m = re.match(r'\s*\.section\s+\.text\.(\S+?),', ln)
        if m: cur_fn = m.group(1); cur_line = None; continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m: cur_line = (os.path.basename(m.group(1)), int(m.group(2)), m.group(3)); continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
        if m and cur_fn: line_of[(cur_fn, int(m.group(1), 16))] = cur_line
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kre], capture_output=True, text=True).stdout
    blocks, cur = [], None
    for r in csv.reader(out.splitlines()):
        if r and r[0] == 'Kernel Name': cur = []; blocks.append((r[1], cur))
        elif cur is not None: cur.append(r)
    name, b = blocks[int(idx)]
    h, data = b[0], b[1:]
    iA, iI, iSm, iS = h.index('Address'), h.index('Instructions Executed'), h.index('# Samples'), h.index('Source')
    base = int(data[0][iA], 16)
    fn = [f for (f, _) in line_of if fnpat in f][0]
    rows = [(int(r[iSm]), int(r[iI]), int(r[iA], 16) - base, r[iS]) for r in data]
    tots = sum(r[0] for r in rows)
    print(name[:70], 'samples', tots)
    for sm, ex, off, src in sorted(rows, reverse=True)[:int(top)]:
        print(f'{100*sm/tots:5.1f}% samp {ex:9d} exec  +{off:05x}  {src[:60]:60s} {line_of.get((fn, off))}')

if __name__ == '__main__':
    main(*sys.argv[1:])
