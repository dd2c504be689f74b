#!/usr/bin/env python
"""Throughput of the bench workload against the number of instance groups, device-driven vs host-polled solve.

    python tools/groups_sweep.py [steps]        # on the GPU box; prints one line per setting
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
steps = sys.argv[1] if len(sys.argv) > 1 else "15"
for host_loop, unroll in (("0", "0"), ("0", "10"), ("0", "13"), ("1", "0")):
    for g in (1, 2, 4, 8):
        if host_loop == "1" and g > 1:
            continue                      # host-polled groups run one after the other since the worker threads are gone
        env = dict(os.environ, MPCB_GROUPS=str(g), MPCB_HOST_LOOP=host_loop, MPCB_GRAPH_UNROLL=unroll, MPCB_BENCH_NOSAMPLER="1")
        res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", steps, "--warmup", "4", "--no-cpu-baseline"],
                             env=env, capture_output=True, text=True)
        try:
            d = json.loads(res.stdout.strip().splitlines()[-1])
            print("host_loop=%s unroll=%2s groups=%2d value %9.0f e2e %9.0f ms/step %6.2f p50 %6.2f launches %d" % (
                host_loop, unroll, g, d["value"], d["e2e"]["value"], d["ms_per_step"], d["p50_step_latency_ms"], d["gpu_launches"]), flush=True)
        except Exception as exc:
            print(host_loop, g, "FAILED", exc, res.stderr[-600:], flush=True)
