"""Per-kernel SASS opcode histogram of libpb2.so (cuobjdump -sass): the tensor-core / TMEM / TMA / async-copy / MUFU /
barrier mnemonics that show which hardware path each kernel uses.  Writes profiles/r02_sass_histogram.md."""
import collections
import re
import subprocess
import sys

LIB = 'probability_b200/_C/libpb2.so'
OPS = ['UTCHMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UBLKCP', 'LDGSTS', 'SYNCS', 'MUFU', 'BAR', 'HMMA', 'FFMA', 'LDL', 'STL',
       'USETMAXREG']


def main():
  out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
  per = collections.OrderedDict()
  name = None
  for line in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
      name = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
      name = re.sub(r'\(.*', '', name).replace('pb2::', '').replace('(anonymous namespace)::', '')
      per[name] = collections.Counter()
      continue
    m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
    if m and name:
      op = m.group(1)
      per[name]['_total'] += 1
      for o in OPS:
        if op == o or op.startswith(o + '.') or (o == 'BAR' and op in ('BAR', 'BAR.SYNC', 'BAR.RED', 'BAR.ARV')):
          per[name][o] += 1
  rows = [(n, c) for n, c in per.items() if c['_total'] > 0]
  with open('profiles/r02_sass_histogram.md', 'w') as f:
    f.write('# SASS opcode histogram per kernel of `libpb2.so` (sm_100a; `python scripts/sass_histogram.py`)\n\n')
    f.write('Static instruction counts.  `UTCHMMA` = tcgen05.mma, `UTCBAR` = tcgen05.commit, `LDTM`/`STTM` = tcgen05.ld/st (TMEM), '
            '`UTMALDG` = cp.async.bulk.tensor (TMA tensor load), `LDGSTS` = cp.async, `SYNCS` = mbarrier ops, `MUFU` = special-function '
            'unit, `LDL`/`STL` = local-memory (spill) accesses, `USETMAXREG` = setmaxnreg.\n\n')
    f.write('| kernel | instr | ' + ' | '.join(OPS) + ' |\n|---|---|' + '---|' * len(OPS) + '\n')
    for n, c in sorted(rows, key=lambda r: -r[1]['UTCHMMA'] * 100000 - r[1]['_total']):
      short = n if len(n) < 110 else n[:107] + '...'
      f.write('| `%s` | %d | ' % (short, c['_total']) + ' | '.join(str(c[o]) if c[o] else '' for o in OPS) + ' |\n')
    tot = collections.Counter()
    for _, c in rows:
      tot.update(c)
    f.write('| **all kernels** | %d | ' % tot['_total'] + ' | '.join(str(tot[o]) for o in OPS) + ' |\n')
  print(open('profiles/r02_sass_histogram.md').read()[:3000])


if __name__ == '__main__':
  main()
