#!/bin/bash
# Scaling run on an N-GPU box:  bash tools/scale_run.sh N [steps]   (writes gpurun_out/r02_scale_<N>gpu.json and the sharded ENMPC line)
N=${1:-8}; K=${2:-40}
mkdir -p gpurun_out
if [ "$N" = "1" ]; then
  python bench.py --gpus 1 --steps $K --warmup 5 --no-cpu-baseline > gpurun_out/r02_scale_${N}gpu.json 2> gpurun_out/r02_scale_${N}gpu.err
  MPCB_ENMPC_TOTAL=4096 python tools/enmpc_sharded.py > gpurun_out/r02_enmpc_${N}gpu.json 2> gpurun_out/r02_enmpc_${N}gpu.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps $K --warmup 5 > gpurun_out/r02_scale_${N}gpu.json 2> gpurun_out/r02_scale_${N}gpu.err
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/enmpc_sharded.py > gpurun_out/r02_enmpc_${N}gpu.json 2> gpurun_out/r02_enmpc_${N}gpu.err
fi
tail -c 600 gpurun_out/r02_scale_${N}gpu.json; echo; cat gpurun_out/r02_enmpc_${N}gpu.json; tail -3 gpurun_out/r02_enmpc_${N}gpu.err
