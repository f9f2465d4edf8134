#!/usr/bin/env python
"""Static instruction counts per kernel of the shipped library (cuobjdump -sass; runs without a GPU).
Regenerated from the built .so by __graft_entry__.build() callers / by hand:
    python scripts/sass_evidence.py [out.txt]"""
import os, re, subprocess, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, 'digicamtoy_b200', 'libdigicamtoy_sm100.so')
out = open(sys.argv[1], 'w') if len(sys.argv) > 1 else sys.stdout
elf = subprocess.run(['cuobjdump', '-lelf', so], capture_output=True, text=True).stdout
sass = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
MNEMONICS = ['UTCHMMA', 'UTCBAR', 'UTCATOMSWS', 'SYNCS', 'LDTM', 'STTM', 'FFMA2', 'FFMA', 'HFMA2', 'F2FP.F16.F32.PACK_AB', 'LDS.128', 'LDS', 'STS.128',
             'STS.64', 'STS', 'ATOMS', 'RED', 'REDUX', 'CREDUX', 'MUFU.LG2', 'MUFU.SQRT', 'MUFU.COS', 'MUFU.SIN', 'IMAD.WIDE.U32',
             'STG.E.EF.128', 'LDGSTS', 'PRMT', 'NANOSLEEP', 'DFMA', 'STL', 'LDL']
out.write('cuobjdump of digicamtoy_b200/libdigicamtoy_sm100.so -- static instruction counts per kernel (prefix match on the mnemonic)\n')
out.write('ELF: ' + ' '.join(l.strip() for l in elf.splitlines() if 'sm_' in l) + '\n')
out.write('UTCHMMA = tcgen05.mma (5th-generation tensor core, accumulator in TMEM), UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st,\n'
          'SYNCS = mbarrier, FFMA2 = packed FP32 FMA (Blackwell), LDGSTS = cp.async staging of the statistics tiles, STG.E.EF.128 = 16-byte\n'
          'streaming stores of the cube, STL / LDL = local-memory traffic (spills and by-reference arguments of out-of-line samplers).\n\n')
cur, counts, total = None, None, 0
def flush():
    if cur:
        name = subprocess.run(['c++filt', cur], capture_output=True, text=True).stdout.strip() or cur
        out.write('%s   total %d\n   %s\n' % (name, total, ', '.join('%s=%d' % (k, v) for k, v in counts.items() if v)))
for line in sass.splitlines():
    m = re.search(r'Function : (\S+)', line)
    if m:
        flush()
        cur, counts, total = m.group(1), collections.OrderedDict((k, 0) for k in MNEMONICS), 0
        continue
    m = re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)', line)
    if m and cur:
        total += 1
        op = m.group(1)
        for k in MNEMONICS:
            if op == k or op.startswith(k + '.'):
                counts[k] += 1
                break
flush()
