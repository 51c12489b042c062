"""Regenerate profiles/sass/sass_evidence.txt: Blackwell-specific SASS mnemonics per kernel of _b200_ops.so."""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else 'coinstac_dinunet_b200/ops/_b200_ops.so'
out = sys.argv[2] if len(sys.argv) > 2 else 'profiles/sass/sass_evidence.txt'
sass = subprocess.run(['cuobjdump', '-sass', so], capture_output=True, text=True).stdout
names = subprocess.run(['c++filt'], input='\n'.join(re.findall(r'Function : (\S+)', sass)), capture_output=True, text=True).stdout.split('\n')
parts = re.split(r'\n\s*Function : \S+', sass)[1:]
KEY = re.compile(r'^(UTC\w*|UTMA\w*|UBLKCP|LDTM|STTM|SYNCS|MULTIMEM|LDGMC|LDGSTS|REDG|ATOMS|ATOMG|FENCE|MEMBAR|REDUX|ELECT|R2UR\.BROADCAST|STG\.E\.ENL2\.256)')
lines = ['# SASS evidence per kernel (cuobjdump -sass coinstac_dinunet_b200/ops/_b200_ops.so, sm_100a); scripts/sass_evidence.py',
         '# tcgen05.mma -> UTCHMMA ; tcgen05.ld -> LDTM ; TMA -> UTMALDG ; tcgen05.commit -> UTCBAR ; mbarrier -> SYNCS ;',
         '# multimem.* -> MULTIMEM / LDGMC ; cp.async -> LDGSTS ; 256-bit store -> STG.E.ENL2.256 ;',
         '# R2UR.BROADCAST + ELECT around a UTCHMMA = the non-uniform "waterfall" issue path (0 in the hot kernels since the',
         '# warp-uniform issuer rewrite)', '']
for name, body in zip(names, parts):
    ops = collections.Counter()
    total = 0
    for m in re.finditer(r'/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z][A-Z0-9_.]*)', body):
        total += 1
        op = m.group(1)
        if KEY.match(op):
            ops['.'.join(op.split('.')[:4])] += 1
    if not any(k.startswith(('UTC', 'UTMA', 'MULTIMEM', 'LDGMC', 'STG.E.ENL2.256')) for k in ops):
        continue
    lines.append(name)
    lines.append(f'   total instr: {total}')
    lines.append('   ' + ', '.join(f'{k} x{v}' for k, v in sorted(ops.items())))
    lines.append('')
open(out, 'w').write('\n'.join(lines))
print(out, len(lines))
