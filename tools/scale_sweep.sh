#!/bin/bash
# 1 -> 8 GPU weak-scaling sweep on one 8-GPU box, launched the way the driver launches it:
#   bash tools/scale_sweep.sh [steps] [tag]     (writes gpurun_out/<tag>_scale_<N>gpu.json)
K=${1:-40}; TAG=${2:-r02_v7}
mkdir -p gpurun_out
for N in 1 2 4 8; do
  if [ "$N" = "1" ]; then
    python bench.py --gpus 1 --steps $K --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_scale_${N}gpu.json 2> gpurun_out/${TAG}_scale_${N}gpu.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520 + N)) \
        bench.py --gpus $N --steps $K --warmup 5 > gpurun_out/${TAG}_scale_${N}gpu.json 2> gpurun_out/${TAG}_scale_${N}gpu.err
  fi
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_scale_${N}gpu.json").read().strip().splitlines()[-1])
    print("N=%d value %.0f e2e %.0f ms/step %.2f clocks %s" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"], d["clocks"]))
except Exception as exc:
    print("N=${N} FAILED", exc)
PY
done
# the reference arm under torchrun: rank 0 alone works and prints
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/${TAG}_reference_torchrun2.json 2> gpurun_out/${TAG}_reference_torchrun2.err
wc -l gpurun_out/${TAG}_reference_torchrun2.json; cut -c1-200 gpurun_out/${TAG}_reference_torchrun2.json
