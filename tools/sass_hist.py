"""Opcode histogram of the shipped library's SASS (cuobjdump -sass), per kernel and overall.
Usage: python tools/sass_hist.py [libgnxb200.so] > profiles/rNN_sass_opcodes.txt"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else 'geonomics_b200/libgnxb200.so'
txt = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
ins = re.compile(r'^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_.]+)?)')
fn = None
per = collections.OrderedDict()
for line in txt.splitlines():
    if 'Function :' in line:
        fn = line.split('Function :')[1].strip()
        per[fn] = collections.Counter()
        continue
    m = ins.match(line)
    if m and fn:
        per[fn][m.group(1).split('.')[0]] += 1
tot = collections.Counter()
for c in per.values():
    tot.update(c)
demangle = subprocess.run(['c++filt'], input='\n'.join(per), capture_output=True, text=True).stdout.splitlines()
print('# SASS opcode histogram of %s (sm_100a), %d kernels, %d instructions' % (lib, len(per), sum(tot.values())))
print('# bulk-copy / TMA / tensor-core mnemonics: ' + ', '.join(
    '%s=%d' % (k, sum(v for op, v in tot.items() if op.startswith(k)))
    for k in ('UBLKCP', 'UTMALDG', 'UTMASTG', 'SYNCS', 'UTCMMA', 'HMMA', 'DFMA', 'DADD', 'DMUL', 'LDG', 'STG', 'ATOMG',
              'REDG', 'RED', 'ATOM', 'MATCH', 'VOTE', 'SHFL', 'POPC', 'LOP3')))
print('\n## all kernels')
for op, n in tot.most_common(60):
    print('%9d  %s' % (n, op))
print('\n## per kernel: instructions, then the ten most frequent opcodes')
for (f, c), name in zip(per.items(), demangle):
    name = re.sub(r'\(.*', '', name)
    print('%7d  %-48s %s' % (sum(c.values()), name[:48], ' '.join('%s:%d' % kv for kv in c.most_common(10))))
