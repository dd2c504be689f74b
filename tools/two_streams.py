#!/usr/bin/env python
"""Experiment: drive two half-batches from two host threads on two CUDA streams (latency-bound kernels of one half
overlap the throughput-bound evaluation of the other).  python tools/two_streams.py [groups] [steps]"""
import os, sys, threading, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from mpc_code_b200.mpc_loop import CompiledProblem

G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
STAGGER_MS = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0      # group g starts g * STAGGER_MS late (phase offset)
B = bench.BATCH_PER_GPU
prob, ss, ocp = bench._problem()
cp = CompiledProblem(prob, "nmpc_cstr")
x0, noise = bench._workload(prob, B, K + 5)
dev = torch.device("cuda", 0)
per = B // G
ctls, streams, noises = [], [], []
for g in range(G):
    st = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(st):
        c = cp.controller(per); c.reset(x0_p=x0[g * per:(g + 1) * per], x0_m=x0[g * per:(g + 1) * per])
        noises.append(torch.as_tensor(noise[:, g * per:(g + 1) * per, :], device=dev))
    ctls.append(c); streams.append(st)
torch.cuda.synchronize()

def run(g, k0, k1):
    if STAGGER_MS > 0:
        time.sleep(1e-3 * STAGGER_MS * g)
    with torch.cuda.stream(streams[g]):
        for k in range(k0, k1):
            ctls[g].step_fused(noises[g][k])
        streams[g].synchronize()

for phase, (k0, k1) in (("warmup", (0, 5)), ("timed", (5, 5 + K))):
    torch.cuda.synchronize(); t0 = time.time()
    th = [threading.Thread(target=run, args=(g, k0, k1)) for g in range(G)]
    [t.start() for t in th]; [t.join() for t in th]
    torch.cuda.synchronize(); dt = time.time() - t0
    print(phase, "groups", G, "stagger_ms", STAGGER_MS, "steps/s %.0f  ms/step %.2f" % (B * (k1 - k0) / dt, 1e3 * dt / (k1 - k0)), flush=True)
