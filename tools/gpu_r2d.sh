#!/bin/bash
# Usage: gpurun --gpus 2 --timeout 1800 -- bash tools/gpu_r2d.sh <tag> <ngpu>
TAG=${1:-r2d}; NG=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt; nvidia-smi topo -m >> $OUT/gpus.txt 2>&1
echo "== pytest multi"; timeout 900 python -m pytest tests/test_gpu_multi.py "tests/test_gpu_features.py::test_plan_cache_reuses_nodes_and_stays_correct" -m gpu -q --durations=5 2>&1 | tee $OUT/pytest_multi.log | tail -15
for RED in nccl peer; do
  echo "== bench weak N=$NG reduce=$RED"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $NG --steps 10 --warmup 3 --reduce $RED 2>$OUT/bench_weak_$RED.err | tee $OUT/bench_weak_$RED.log | tail -1 | cut -c1-400
done
echo "== bench strong cfg3 N=$NG"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $NG --steps 10 --warmup 3 --scaling strong 2>$OUT/bench_strong.err | tee $OUT/bench_strong.log | tail -1 | cut -c1-400
echo "== group (C plan API, one process) cfg3"
timeout 600 python tools/bench_group.py --config cfg3 --devices 1,$NG --check 2>&1 | tee $OUT/group_cfg3.log | cut -c1-600
ls -la $OUT
