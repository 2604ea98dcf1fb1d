#!/usr/bin/env python
"""Per-source-line instruction / stall-sample totals of one kernel in an .ncu-rep.

ncu's `--page source --csv` lists SASS with counters but no source lines; `nvdisasm -g` lists
SASS with source lines but no counters.  This joins the two by instruction offset.

  python tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [--lib geonomics_b200/libgnxb200.so] [--top 40]

The library must be the build that was profiled (same SASS for that kernel).
"""
import argparse
import csv
import io
import os
import re
import subprocess
import tempfile


def disasm_lines(lib, kernel_re):
    tmp = tempfile.mkdtemp()
    subprocess.run(['cuobjdump', '-xelf', 'all', os.path.abspath(lib)], cwd=tmp, check=True,
                   stdout=subprocess.DEVNULL)
    cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith('.cubin')][0]
    txt = subprocess.run(['nvdisasm', '-g', '-c', cubin], check=True, capture_output=True, text=True).stdout
    funcs = {}
    cur = None
    line = None
    for ln in txt.splitlines():
        m = re.match(r'\s*\.global\s+(\S+)', ln)
        if m:
            cur = m.group(1)
            funcs[cur] = {}
            line = None
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*);', ln)
        if m and cur is not None:
            funcs[cur][int(m.group(1), 16)] = (line, m.group(2).strip())
    demangled = subprocess.run(['cu++filt'] + list(funcs), capture_output=True, text=True).stdout.splitlines()
    out = {}
    for mangled, dem in zip(funcs, demangled):
        if re.search(kernel_re, dem):
            out[dem] = funcs[mangled]
    return out


def ncu_source(rep, kernel_re):
    txt = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', 'regex:' + kernel_re],
                         capture_output=True, text=True).stdout
    kernels = []
    rows = list(csv.reader(io.StringIO(txt)))
    i = 0
    while i < len(rows):
        if rows[i] and rows[i][0] == 'Kernel Name':
            name = rows[i][1]
            hdr = rows[i + 1]
            j = i + 2
            body = []
            while j < len(rows) and not (rows[j] and rows[j][0] == 'Kernel Name'):
                if len(rows[j]) == len(hdr):
                    body.append(rows[j])
                j += 1
            kernels.append((name, hdr, body))
            i = j
        else:
            i += 1
    return kernels


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('report')
    ap.add_argument('kernel')
    ap.add_argument('--lib', default='geonomics_b200/libgnxb200.so')
    ap.add_argument('--top', type=int, default=40)
    ap.add_argument('--launch', type=int, default=0)
    a = ap.parse_args()
    dis = disasm_lines(a.lib, a.kernel)
    ks = ncu_source(a.report, a.kernel)
    name, hdr, body = ks[a.launch]
    ia, ie, isamp = hdr.index('Address'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    isrc = hdr.index('Source')
    base = int(body[0][ia], 16) if body[0][ia].startswith('0x') else int(body[0][ia])

    def mismatches(table):
        bad = 0
        for r in body:
            addr = (int(r[ia], 16) if r[ia].startswith('0x') else int(r[ia])) - base
            if table.get(addr, (None, ''))[1].split()[:1] != r[isrc].split()[:1]:
                bad += 1
        return bad
    # several template instantiations can match the regex: take the one whose SASS agrees
    fn = min(dis.values(), key=mismatches)
    per = {}
    tot_e = tot_s = 0
    mism = 0
    for r in body:
        addr = (int(r[ia], 16) if r[ia].startswith('0x') else int(r[ia])) - base
        e, s = int(r[ie] or 0), int(r[isamp] or 0)
        line, sass = fn.get(addr, (None, ''))
        if sass.split()[:1] != r[hdr.index('Source')].split()[:1]:
            mism += 1
        d = per.setdefault(line, [0, 0, 0])
        d[0] += e
        d[1] += s
        d[2] += 1
        tot_e += e
        tot_s += s
    print('%s\n  %d SASS instructions, %d executed (warp-level), %d samples, %d address mismatches'
          % (name, len(body), tot_e, tot_s, mism))
    src_cache = {}
    for line, (e, s, nin) in sorted(per.items(), key=lambda kv: -kv[1][0])[:a.top]:
        text = ''
        if line:
            for root in ('geonomics_b200/csrc', '.'):
                pth = os.path.join(root, line[0])
                if os.path.exists(pth):
                    src_cache.setdefault(pth, open(pth).read().splitlines())
                    if line[1] - 1 < len(src_cache[pth]):
                        text = src_cache[pth][line[1] - 1].strip()
                    break
        print('%5.1f%% inst %5.1f%% samp %4d sass  %s:%s  %s'
              % (100.0 * e / max(1, tot_e), 100.0 * s / max(1, tot_s), nin,
                 line[0] if line else '?', line[1] if line else '?', text[:100]))


if __name__ == '__main__':
    main()
