#!/usr/bin/env python
"""Join an `ncu --page source --csv` SASS listing of one kernel with `nvdisasm -g` line info and
aggregate the warp-stall samples per CUDA source line (the CSV source page carries no line column).

  python tools/ncu_lines.py <report.ncu-rep> <kernel regex> <cubin> <mangled-name substring> [top]

Instructions are matched by their order inside the kernel (same build => same SASS)."""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict


def sass_rows(rep, kregex):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kregex}"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    body = []
    for r in rows[h + 1:]:
        if r and r[0] == "Kernel Name":      # second launch of the same kernel: keep the first only
            break
        if len(r) == len(hdr):
            body.append(r)
    return hdr, body


def line_table(cubin, fn):
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
    start = next(i for i, l in enumerate(txt) if l.startswith(".text.") and fn in l)
    lines, cur = [], None
    for l in txt[start + 1:]:
        if l.startswith("//-----") or l.startswith(".text."):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            lines.append(cur)
    return lines


def main():
    rep, kregex, cubin, fn = sys.argv[1:5]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    hdr, body = sass_rows(rep, kregex)
    lt = line_table(cubin, fn)
    if len(lt) != len(body):
        print(f"warning: {len(body)} profiled instructions vs {len(lt)} disassembled", file=sys.stderr)
    isamp = hdr.index("# Samples")
    stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    per = defaultdict(lambda: [0, defaultdict(int), 0])
    tot = 0
    for k, r in enumerate(body):
        key = lt[k] if k < len(lt) else None
        s = int(r[isamp] or 0)
        per[key][0] += s
        per[key][2] += 1
        tot += s
        for i in stalls:
            v = int(r[i] or 0)
            if v:
                per[key][1][hdr[i]] += v
    print(f"total samples {tot}, instructions {len(body)}")
    for key, (s, st, n) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        why = ", ".join(f"{a.replace('stall_', '')}:{b}" for a, b in sorted(st.items(), key=lambda x: -x[1])[:3])
        print(f"{100.0 * s / max(tot, 1):6.2f}%  {s:6d}  {n:5d} instr  {key}  [{why}]")


if __name__ == "__main__":
    main()
