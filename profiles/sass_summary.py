#!/usr/bin/env python
"""SASS listing summary of the CG-step kernels in the built library (runs without a GPU):

    python profiles/sass_summary.py > profiles/r02b_sass_step_kernels.md

Per kernel: instruction count, opcode histogram of the classes that matter here (global / shared / local memory,
FP64, barriers, async copies), registers are in profiles/*ptxas*.  Presence of UTMALDG / UTMASTG / UBLKCP would prove TMA,
LDGSTS = cp.async, STL / LDL = spills."""
import collections
import re
import subprocess
import sys

LIB = 'ffthompy_b200/libffthom_b200.so'
WANT = ['k_fwd_last_fast<256, 6, 4, 3>', 'k_fwd_last_fast<256, 6, 4, 1>', 'k_c2c_fast<256, 8, false>', 'k_c2c_fast<256, 8, true>',
        'k_mid_green_pipe<256, 4, 1, 3>', 'k_mid2<256, 1, 1, 0>', 'k_inv_last_fast<256, 6, 4, 1>', 'k_cg_update_r(',
        'k_fwd_last_reg3<512, 6, 2, 3>', 'k_c2c_reg3_map<512, 8, false>', 'k_mid_green_reg3<512, 4, 1, 3, 768>',
        'k_inv_last_reg3<512, 6, 2>', 'k_mid_green_512<1>', 'k_mid3<1>', 'k_fwd_last_odd<255, 3, 8, 1, 2>',
        'k_c2c_fast<255, 8, false>', 'k_mid_green_odd<255, 4, 0, 2>', 'k_inv_last_odd<255, 3, 8>', 'k_cg_update_r1(',
        'k_assemble_AH<6, 6>', 'k_topologies']
CLASSES = collections.OrderedDict([
    ('LDG', r'^LDG'), ('STG', r'^STG'), ('LDS', r'^LDS'), ('STS', r'^STS'), ('LDGSTS (cp.async)', r'^LDGSTS'),
    ('LDL/STL (spills)', r'^(LDL|STL)'), ('DFMA', r'^DFMA'), ('DADD', r'^DADD'), ('DMUL', r'^DMUL'), ('BAR', r'^BAR'),
    ('TMA (UTMALDG/UTMASTG/UBLKCP)', r'^(UTMALDG|UTMASTG|UBLKCP)'), ('SHFL', r'^SHFL')])


def main():
    txt = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    blocks = re.split(r'\n\s*Function : ', txt)[1:]
    names = subprocess.run(['c++filt'], input='\n'.join(b.split('\n', 1)[0].strip() for b in blocks), capture_output=True,
                           text=True).stdout.split('\n')
    print('# SASS of the step kernels (`cuobjdump -sass %s`, sm_100a)\n' % LIB)
    print('| kernel | instructions | ' + ' | '.join(CLASSES) + ' |')
    print('|---|---|' + '---|'*len(CLASSES))
    for want in WANT:
        for name, blk in zip(names, blocks):
            if want in name:
                ops = re.findall(r'/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)', blk)
                hist = [sum(1 for o in ops if re.match(pat, o)) for pat in CLASSES.values()]
                print('| `%s` | %d | %s |' % (want, len(ops), ' | '.join(str(h) for h in hist)))
                break
    print('\nNo kernel of the library uses TMA; the asynchronous path of the axis-0 pass is per-thread 16-byte cp.async '
          '(`LDGSTS`).  `k_mid2` (opt-in, FH_MID2=1) carries the local-memory traffic of its in-register Green stage (register '
          'spills), one reason it lost against `k_mid_green_pipe` (DESIGN.md section 4); the few LDL/STL of the 512 and odd-length '
          'kernels come from their register caps (three / two CTAs per SM).')


if __name__ == '__main__':
    sys.exit(main())
