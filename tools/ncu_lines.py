#!/usr/bin/env python
"""Per-source-line stall samples / instruction counts of an ncu report captured with --import-source on.
    python tools/ncu_lines.py gpurun_out/prof_kkt6.ncu-rep [top]"""
import collections, csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
kernels = []
agg = None
cur, hdr = None, None
def num(v):
    try: return int(float(v))
    except Exception: return 0
seen_kernel = 0
for r in rows:
    if len(r) == 2 and r[0] == "Function Name":
        agg = collections.defaultdict(lambda: [0, 0, 0, ""]); kernels.append(agg)
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = r; wi = hdr.index("Warp Stall Sampling (All Samples)"); ii = hdr.index("Instructions Executed"); li = hdr.index("stall_long_sb"); continue
    if hdr and len(r) > li:
        try: key = (cur, int(r[0]))
        except Exception: continue
        a = agg[key]; a[0] += num(r[wi]); a[1] += num(r[ii]); a[2] += num(r[li]); a[3] = r[1]
agg = max(kernels, key=lambda g: sum(a[1] for a in g.values()))      # the busiest captured launch
tot = sum(a[0] for a in agg.values()) or 1; toti = sum(a[1] for a in agg.values()) or 1
print("# total samples %d, total warp instructions %d" % (tot, toti))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 30]:
    print("%5.1f%% inst %5.1f%% smp longsb %5d  %s:%d  %s" % (100 * a[1] / toti, 100 * a[0] / tot, a[2], key[0], key[1], a[3].strip()[:105]))
