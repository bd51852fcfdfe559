"""Per-CUDA-source-line instruction counts and stall samples from an .ncu-rep captured with --import-source on.

    python tools/ncu_lines.py report.ncu-rep [top_n]
"""
import csv, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur, hdr, lines = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur = r[1].split('/')[-1]
    elif r[0] == 'Line No':
        hdr = r
    elif hdr and len(r) > 8 and r[2] == '-':
        try:
            lines.append((int(r[hdr.index('Instructions Executed')]), int(r[hdr.index('# Samples')]), cur, r[0], r[1].strip()[:110]))
        except ValueError:
            pass
tot = sum(l[0] for l in lines) or 1
tots = sum(l[1] for l in lines) or 1
print(f'total warp instructions {tot}, samples {tots}')
for e, s, f, ln, src in sorted(lines, reverse=True)[:top]:
    print(f'{e:12d} {e/tot:6.3f} smp {s/tots:6.3f}  {f}:{ln:>4s}  {src}')
