"""Per-source-line stall samples of one kernel from an ncu report (needs -lineinfo and --import-source on).
Usage: python tools/ncu_lines.py report.ncu-rep <kernel-substring> [top]"""
import csv
import io
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
kern, fname, hdr, rows = None, None, None, []
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] in ('Kernel Name', 'Function Name'):
        kern = r[1]
    elif r[0] in ('File Name', 'File Path'):
        fname = r[1]
    elif r[0] == 'Line No':
        hdr = r
    elif hdr and r[0] and kern and pat in kern and len(r) >= len(hdr) - 2:
        try:
            rows.append((int(r[hdr.index('# Samples')]), int(r[hdr.index('Instructions Executed')]), fname.split('/')[-1], r[0], r[1].strip()[:100],
                         {k: int(r[i]) for i, k in enumerate(hdr) if k.startswith('stall_') and 'Not' not in k and r[i].isdigit()}))
        except (ValueError, IndexError):
            pass
tot = sum(r[0] for r in rows) or 1
print(f'kernel ~ {pat}: {tot} samples')
stall = {}
for r in rows:
    for k, v in r[5].items():
        stall[k] = stall.get(k, 0) + v
print('stall totals:', ', '.join(f'{k[6:]} {100 * v / tot:.0f}%' for k, v in sorted(stall.items(), key=lambda x: -x[1])[:9]))
for s, ins, f, ln, src, st in sorted(rows, key=lambda x: -x[0])[:top]:
    main = max(st.items(), key=lambda x: x[1])[0][6:] if st else ''
    print(f'{100 * s / tot:5.1f}% inst {ins:9d} {f}:{ln:>4s} [{main:12s}] {src}')
