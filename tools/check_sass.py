"""Build-time guard: the hot loop (layer 2 of the MLP) of every default fused-kernel variant that keeps its
weights in the constant bank must fetch them with LDCU (uniform registers), not per-lane LDC — ptxas decides
this heuristically (see vnet_kernels.cu).  Also reports FFMA2 density of the loop.
Usage: python tools/check_sass.py   (needs cuobjdump; no GPU)"""
import collections
import os
import re
import subprocess
import sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'meta-viterbinet_b200', 'libmvn_b200.so')
DEFAULTS = {1: (1, 448), 2: (1, 448), 3: (1, 448), 4: (1, 384), 5: (1, 384)}   # the FMA variant of every L <= 5


def hot_loop(lines):
    ins = []
    for l in lines:
        m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    best = None
    for a, t in ins:
        m = re.search(r'BRA\S*\s+.*?(0x[0-9a-f]+)', t)
        if m and int(m.group(1), 16) < a:
            body = [x for x in ins if int(m.group(1), 16) <= x[0] <= a]
            nf = sum('FFMA2' in x[1] for x in body)
            if nf >= 50 and (best is None or len(body) < len(best)):
                best = body
    return best


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True).stdout
    funcs = re.split(r'\n\s*Function : ', sass)
    ok = True
    for L, (ws, nt) in DEFAULTS.items():
        pat = f'vnet_decode_kernelILi{L}ENS_12FusedVariantILi{L}ELi2ELi{ws}ELi{nt}ELi10'
        body = next((f for f in funcs if f.startswith('_ZN3mvn18' + pat)), None)
        if body is None:
            print(f'L={L}: default variant {pat} not found in {LIB}')
            ok = False
            continue
        loop = hot_loop(body.splitlines())
        c = collections.Counter(re.sub(r'^@!?U?P\d+\s+', '', t).split()[0].split('.')[0] for _, t in loop)
        good = c['LDCU'] >= 20 * (c['LDC'] + 1) // 2 and c['LDCU'] > c['LDC']
        print(f'L={L} NT={nt}: loop {len(loop)} instrs, FFMA2 {c["FFMA2"]}, LDCU {c["LDCU"]}, LDC {c["LDC"]}, '
              f'MOV {c["MOV"] + c["IMAD"]} -> {"ok" if good else "FALLBACK TO LDC"}')
        ok &= good
    # tcgen05 kernel (default for every L): tensor-core and TMEM instructions present, MMAs issued back to back
    # (elect.sync form: no per-MMA ELECT / R2UR / BRA.U.ANY serialisation loop)
    for L in range(1, 9):
        body = next((f for f in funcs if f.startswith(f'_ZN3mvn21vnet_decode_tc_kernelILi{L}ELb0E')), None)
        if body is None:
            print(f'tcgen05 L={L}: kernel not found')
            ok = False
            continue
        c = collections.Counter(m.group(1) for m in re.finditer(r'/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', body))
        good = c['UTCHMMA'] >= 22 and c['LDTM'] > 0 and c['STTM'] > 0 and c['UTCBAR'] >= 2 and 'BRA.U.ANY' not in body
        print(f'tcgen05 L={L}: UTCHMMA {c["UTCHMMA"]}, LDTM {c["LDTM"]}, STTM {c["STTM"]}, UTCBAR {c["UTCBAR"]}, '
              f'FFMA2 {c["FFMA2"]}, MUFU {c["MUFU"]}, serialised-issue loops {body.count("BRA.U.ANY")} -> {"ok" if good else "CHECK"}')
        ok &= good
    return 0 if ok else 1


if __name__ == '__main__':
    sys.exit(main())
