"""Aggregate an ncu SASS source page per CUDA source line.

    python profiles/ncu_lines.py <report.ncu-rep> <cubin with -lineinfo> [top]

ncu's CSV source page is SASS-only; nvdisasm -g gives the line of every SASS instruction of the same
cubin in the same order, so the two are zipped by instruction index.
"""
import collections
import csv
import re
import subprocess
import sys


def sass_lines(cubin, kernel_substr):
    txt = subprocess.run(['nvdisasm', '-g', '-c', cubin], capture_output=True, text=True).stdout
    out, cur_line, in_fn = [], None, False
    for ln in txt.splitlines():
        if ln.startswith('.text.') or re.match(r'\s*\.section\s+\.text\.', ln):
            in_fn = kernel_substr in ln
        if not in_fn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur_line = (m.group(1).split('/')[-1], int(m.group(2)))
            continue
        if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', ln):
            out.append(cur_line)
    return out


def main(rep, cubin, top=40, kernel='ie_resolve_tile_kernel'):
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    hdr = rows[hdr_i]
    body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
    lines = sass_lines(cubin, kernel)
    print(f"# {len(body)} SASS rows, {len(lines)} disassembled instructions")
    ci = {h: i for i, h in enumerate(hdr)}
    agg = collections.defaultdict(lambda: [0, 0, 0, 0])
    n = min(len(body), len(lines))
    for k in range(n):
        r = body[k]
        a = agg[lines[k]]
        a[0] += int(r[ci['Instructions Executed']] or 0)
        a[1] += int(r[ci['Thread Instructions Executed']] or 0)
        a[2] += int(r[ci['# Samples']] or 0)
        a[3] += int(r[ci['stall_long_sb']] or 0)
    tot_i = sum(a[0] for a in agg.values()); tot_s = sum(a[2] for a in agg.values())
    print(f"# total warp instr {tot_i}, samples {tot_s}")
    for key, a in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
        print(f"{str(key):32s} inst {a[0]:>11d} ({100*a[0]/tot_i:5.1f}%) lanes {a[1]/max(a[0],1):5.1f} samples {a[2]:>7d} ({100*a[2]/tot_s:5.1f}%) long_sb {a[3]}")


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 40, sys.argv[4] if len(sys.argv) > 4 else 'ie_resolve_tile_kernel')


def phase_table(rep, cubin, ranges, kernel='ie_resolve_tile_kernel'):
    """ranges: list of (name, file, lo, hi) -> share of warp instructions and samples."""
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
    hdr = rows[hdr_i]
    body = [r for r in rows[hdr_i + 1:] if len(r) == len(hdr)]
    lines = sass_lines(cubin, kernel)
    ci = {h: i for i, h in enumerate(hdr)}
    acc = collections.defaultdict(lambda: [0, 0, 0])
    for k in range(min(len(body), len(lines))):
        key = lines[k]
        name = 'other'
        if key:
            for nm, f, lo, hi in ranges:
                if key[0] == f and lo <= key[1] <= hi:
                    name = nm
                    break
            else:
                name = key[0]
        a = acc[name]
        a[0] += int(body[k][ci['Instructions Executed']] or 0)
        a[1] += int(body[k][ci['Thread Instructions Executed']] or 0)
        a[2] += int(body[k][ci['# Samples']] or 0)
    ti = sum(a[0] for a in acc.values()); ts = sum(a[2] for a in acc.values())
    for nm, a in sorted(acc.items(), key=lambda kv: -kv[1][0]):
        print(f"{nm:28s} inst {a[0]:>11d} ({100*a[0]/ti:5.1f}%)  lanes {a[1]/max(a[0],1):5.1f}  samples {100*a[2]/ts:5.1f}%")
