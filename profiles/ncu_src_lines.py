"""Per-CUDA-source-line instruction shares of one ncu capture (taken with --import-source on).

    python profiles/ncu_src_lines.py <report.ncu-rep> [min_pct]

Uses ncu's own `--page source --print-source cuda,sass` view, so no matching cubin is needed.
"""
import csv
import subprocess
import sys


def load(rep):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                         capture_output=True, text=True).stdout
    cur, hdr, rows = None, None, []
    for r in csv.reader(raw.splitlines()):
        if not r:
            continue
        if r[0] == 'File Path':
            cur = r[1].split('/')[-1]
        elif r[0] == 'Line No':
            hdr = {h: i for i, h in enumerate(r)}
        elif hdr and r[0].isdigit() and len(r) > hdr['# Samples']:
            try:
                rows.append((cur, int(r[0]), r[1].strip()[:100], int(r[hdr['Instructions Executed']] or 0),
                             int(r[hdr['Thread Instructions Executed']] or 0), int(r[hdr['# Samples']] or 0)))
            except ValueError:
                pass
    return rows


def main():
    rows = load(sys.argv[1])
    min_pct = float(sys.argv[2]) if len(sys.argv) > 2 else 0.2
    tot = sum(x[3] for x in rows) or 1
    ts = sum(x[5] for x in rows) or 1
    print(f'# warp instructions {tot}, samples {ts}')
    for f, l, s, i, t, sm in rows:
        if 100 * i / tot >= min_pct or 100 * sm / ts >= 2 * min_pct:
            print(f'{f[:20]:20s}{l:5d} {100 * i / tot:5.2f}% lanes {t / max(i, 1):4.1f} samples {100 * sm / ts:4.1f}% | {s}')


if __name__ == '__main__':
    main()
