#!/usr/bin/env python3
"""Join an ncu report's per-SASS-instruction samples with nvdisasm line info.

  python tools/ncu_lines.py <report.ncu-rep> <cubin> <kernel-substring> [top]

Prints (a) the stall-reason totals, (b) an opcode histogram of executed warp instructions and
(c) the source lines of the kernel ranked by stall samples with their dominant stall reasons.
The cubin must come from the same build as the report (cuobjdump -xelf all lib.so).
"""
import collections
import csv
import io
import re
import subprocess
import sys


def sass_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if "# Samples" in r)
    return rows[hi], [r for r in rows[hi + 1:] if len(r) == len(rows[hi])]


def line_map(cubin, kernel):
    """list of (opcode text, file line) in program order for the first function whose name matches"""
    txt = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
    res, on, line = [], False, 0
    for ln in txt.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
        if m:
            on = kernel in m.group(1)
            continue
        if not on:
            continue
        m = re.search(r'//## File ".*?([^/"]+)", line (\d+)', ln)
        if m:
            line = (m.group(1), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m:
            res.append((m.group(2).strip(), line))
    return res


def main():
    rep, cubin, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    hdr, rows = sass_rows(rep)
    lm = line_map(cubin, kernel)
    si, ie, so = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    if len(lm) != len(rows):
        print("warning: %d SASS rows in report vs %d in cubin" % (len(rows), len(lm)))
    tot = collections.Counter()
    ops = collections.Counter()
    by_line = collections.defaultdict(lambda: [0, 0, collections.Counter()])
    for k, r in enumerate(rows):
        ns, ni = int(r[si] or 0), int(r[ie] or 0)
        op = r[so].split()
        op = (op[1] if op and op[0].startswith("@") else op[0] if op else "?").split(".")[0]
        ops[op] += ni
        line = lm[k][1] if k < len(lm) else ("?", 0)
        e = by_line[line]
        e[0] += ns
        e[1] += ni
        for c in stall_cols:
            v = int(r[c] or 0)
            if v:
                e[2][hdr[c][6:]] += v
                tot[hdr[c][6:]] += v
    nsamp = sum(e[0] for e in by_line.values())
    ninst = sum(ops.values())
    print("samples %d  warp-instructions %d" % (nsamp, ninst))
    print("stalls:", ", ".join("%s %.1f%%" % (k, 100.0 * v / max(1, nsamp)) for k, v in tot.most_common()))
    print("opcodes:", ", ".join("%s %.1f%%" % (k, 100.0 * v / max(1, ninst)) for k, v in ops.most_common(18)))
    print("%-22s %8s %6s %10s  stalls" % ("line", "samples", "%", "instr"))
    for line, e in sorted(by_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-22s %8d %6.1f %10d  %s" % ("%s:%d" % line, e[0], 100.0 * e[0] / max(1, nsamp), e[1],
                                            " ".join("%s=%d" % kv for kv in e[2].most_common(4))))


if __name__ == "__main__":
    main()
