#!/usr/bin/env python
"""Compact per-launch summary of an .ncu-rep (`--set full`): time, DRAM bytes, issue, occupancy, top stalls.

  python tools/ncu_summary.py REPORT.ncu-rep [--csv out.csv]
"""
import argparse
import csv
import io
import subprocess

COLS = [('gpu__time_duration.sum', 'us'), ('dram__bytes_read.sum', 'rd'), ('dram__bytes_write.sum', 'wr'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue%'),
        ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occ%'),
        ('launch__registers_per_thread', 'regs'), ('smsp__inst_executed.sum', 'winst'),
        ('l1tex__t_sector_hit_rate.pct', 'l1hit%'), ('lts__t_sector_hit_rate.pct', 'l2hit%')]
STALLS = ['long_scoreboard', 'short_scoreboard', 'wait', 'math_pipe_throttle', 'lg_throttle', 'not_selected',
          'barrier', 'membar', 'branch_resolving', 'mio_throttle', 'no_instruction', 'dispatch_stall', 'drain',
          'imc_miss', 'tex_throttle', 'sleeping']


def to_bytes(v, unit):
    v = float(v.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}.get(unit, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('report')
    ap.add_argument('--csv')
    a = ap.parse_args()
    txt = subprocess.run(['ncu', '-i', a.report, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    h, units = rows[0], rows[1]
    ix = {n: i for i, n in enumerate(h)}
    out = []
    for r in rows[2:]:
        name = r[ix['Kernel Name']].split('(')[0].replace('void ', '')
        d = {'kernel': name}
        for c, short in COLS:
            if c in ix:
                v = r[ix[c]]
                if short in ('rd', 'wr'):
                    d[short + '_MB'] = round(to_bytes(v, units[ix[c]]) / 1e6, 1)
                elif short == 'us':
                    t = float(v.replace(',', ''))
                    t *= {'ns': 1e-3, 'us': 1, 'ms': 1e3, 's': 1e6}.get(units[ix[c]], 1)
                    d['us'] = round(t, 1)
                else:
                    d[short] = round(float(v.replace(',', '')), 1)
        d["dram_GBs"] = round((d.get("rd_MB", 0) + d.get("wr_MB", 0)) / max(d["us"], 1e-9) * 1e3, 0)
        st = []
        for s in STALLS:
            c = 'smsp__average_warps_issue_stalled_%s_per_issue_active.ratio' % s
            if c in ix:
                st.append((float(r[ix[c]].replace(',', '')), s))
        st.sort(reverse=True)
        d['stalls'] = ' '.join('%s=%.1f' % (s, v) for v, s in st[:4])
        out.append(d)
    keys = ['kernel', 'us', 'rd_MB', 'wr_MB', 'dram_GBs', 'issue%', 'occ%', 'regs', 'winst', 'l1hit%', 'l2hit%', 'stalls']
    for d in out:
        print('  '.join('%s=%s' % (k, d.get(k)) for k in keys))
    if a.csv:
        with open(a.csv, 'w', newline='') as f:
            wr = csv.DictWriter(f, fieldnames=keys)
            wr.writeheader()
            for d in out:
                wr.writerow({k: d.get(k) for k in keys})


if __name__ == '__main__':
    main()
