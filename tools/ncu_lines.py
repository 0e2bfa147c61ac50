#!/usr/bin/env python
"""Per-source-line digest of an ncu report's source page (needs -lineinfo + --import-source on).
usage: ncu_lines.py report.ncu-rep [top_n]   -> source lines ranked by stall samples, with warp instructions, threads per
instruction and long-scoreboard share; plus totals per file."""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; hdr = None; lines = []
def num(d, k):
    try: return int(d.get(k, '0') or 0)
    except ValueError: return 0
for r in rows:
    if not r: continue
    if r[0] in ('File Name', 'File Path'): cur = r[1].split('/')[-1]; continue
    if r[0] == 'Line No':
        hdr = ['Line No', 'Text', 'Address', 'Sass'] + r[4:]; continue
    if hdr is None or cur is None or r[0] == '' or not r[0].isdigit(): continue
    d = dict(zip(hdr, r))  # source rows carry the line's aggregate ('-' in the Address / Sass columns)
    ie = num(d, 'Instructions Executed'); te = num(d, 'Thread Instructions Executed'); ss = num(d, '# Samples')
    if ie == 0 and ss == 0: continue
    lines.append((cur, int(r[0]), r[1].strip(), ie, te, ss, d))
tot_i = sum(l[3] for l in lines); tot_t = sum(l[4] for l in lines); tot_s = sum(l[5] for l in lines)
print(f'total warp instr {tot_i}  thread instr {tot_t}  lanes {tot_t/max(tot_i,1):.2f}  samples {tot_s}')
byfile = collections.defaultdict(lambda: [0, 0, 0])
for l in lines:
    b = byfile[l[0]]; b[0] += l[3]; b[1] += l[4]; b[2] += l[5]
for f, b in byfile.items(): print(f'  {f}: instr {b[0]/tot_i:.3f} lanes {b[1]/max(b[0],1):.1f} samples {b[2]/max(tot_s,1):.3f}')
print('file:line  instr%  lanes  samples%  long_sb% | source')
key = (lambda l: -l[3]) if len(sys.argv) > 3 and sys.argv[3] == 'instr' else (lambda l: -l[5])
for l in sorted(lines, key=key)[:top]:
    d = l[6]; lsb = num(d, 'stall_long_sb')
    print(f'{l[0]}:{l[1]:4d} {100*l[3]/tot_i:5.2f} {l[4]/max(l[3],1):5.1f} {100*l[5]/max(tot_s,1):5.2f} {100*lsb/max(tot_s,1):5.2f} | {l[2][:110]}')
