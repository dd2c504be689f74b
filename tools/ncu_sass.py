#!/usr/bin/env python
"""Per-instruction view of one captured launch (ncu --set full --import-source on): opcode mix by executed warp
instructions, stall-reason totals, and the instructions with the most stall samples.

    python tools/ncu_sass.py gpurun_out/prof.ncu-rep [launch index (default: the one with most samples)] [top N]
"""
import collections
import csv
import subprocess
import sys


def sections(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    secs, cur = [], None
    for row in csv.reader(out.splitlines()):
        if row and row[0] == "Kernel Name":
            cur = dict(name=row[1], hdr=None, rows=[]); secs.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = row
        elif cur is not None and row:
            cur["rows"].append(row)
    return secs


def main():
    secs = sections(sys.argv[1])
    def nsamp(s):
        i = s["hdr"].index("# Samples")
        return sum(int(r[i]) for r in s["rows"])
    idx = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2] != "-" else max(range(len(secs)), key=lambda k: nsamp(secs[k]))
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    s = secs[idx]; h = s["hdr"]
    i_src, i_ex, i_smp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    stall_cols = [(c, h.index(c)) for c in h if c.startswith("stall_") and "Not Issued" not in c]
    mix = collections.Counter(); tot_ex = 0
    stalls = collections.Counter()
    for r in s["rows"]:
        op = r[i_src].split()[0] if r[i_src].split() else "?"
        if op.startswith("@"):
            op = r[i_src].split()[1]
        op = op.split(".")[0]
        ex = int(r[i_ex]); mix[op] += ex; tot_ex += ex
        for c, i in stall_cols:
            stalls[c] += int(r[i])
    print("# %s  launch %d of %d: %d SASS instructions, %d warp instructions executed, %d samples" % (
        s["name"], idx, len(secs), len(s["rows"]), tot_ex, nsamp(s)))
    print("# opcode mix (share of executed warp instructions)")
    for op, n in mix.most_common(18):
        print("  %-10s %6.2f %%" % (op, 100.0 * n / tot_ex))
    tot_st = sum(stalls.values())
    print("# stall samples by reason")
    for c, n in stalls.most_common(10):
        print("  %-24s %6.2f %%" % (c, 100.0 * n / max(tot_st, 1)))
    print("# top instructions by samples")
    rows = sorted(s["rows"], key=lambda r: -int(r[i_smp]))[:top]
    for r in rows:
        why = sorted(((int(r[i]), c) for c, i in stall_cols), reverse=True)[:2]
        print("  %6s  %-60s %s" % (r[i_smp], r[i_src].strip()[:60], " ".join("%s=%d" % (c[6:], n) for n, c in why if n)))


if __name__ == "__main__":
    main()
