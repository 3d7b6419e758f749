#!/bin/bash
# Usage: gpurun --gpus N --timeout 1800 -- bash tools/gpu_r2e.sh <tag> <N>
# cfg4 strong scaling (north_star configs[3]) at N GPUs: process-per-GPU bench (torchrun, device-resident value +
# e2e) and the one-process C path (NFFT_B200_DEVICES); at N >= 2 also the cfg3 weak-scaling bench line; at N = 8 the
# cfg5 plan-per-GPU run and the multi-device tests.
TAG=${1:-r2e}; NG=${2:-1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi -L > $OUT/gpus_$NG.txt; nproc >> $OUT/gpus_$NG.txt; free -g >> $OUT/gpus_$NG.txt
CHECK="--no-check"; [ "$NG" = "8" ] && CHECK=""
run_bench() {  # name, args...
  local name=$1; shift
  if [ "$NG" = "1" ]; then
    timeout 900 python bench.py --gpus 1 "$@" 2>$OUT/$name.err | tee $OUT/$name.log | tail -1 | cut -c1-300
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29521 \
        bench.py --gpus $NG "$@" 2>$OUT/$name.err | tee $OUT/$name.log | tail -1 | cut -c1-300
  fi
}
echo "== cfg4 strong N=$NG"; run_bench cfg4_strong_$NG --config cfg4 --scaling strong --steps 5 --warmup 3 $CHECK
echo "== group cfg4 P=$NG"; timeout 900 python tools/bench_group.py --config cfg4 --devices $NG --steps 3 2>&1 | tee $OUT/group_cfg4_$NG.log | cut -c1-400
if [ "$NG" != "1" ]; then
  echo "== cfg3 weak N=$NG"; run_bench cfg3_weak_$NG --steps 10 --warmup 3
  echo "== cfg3 strong N=$NG"; run_bench cfg3_strong_$NG --scaling strong --steps 10 --warmup 3 --no-check
  echo "== group cfg3 P=$NG"; timeout 600 python tools/bench_group.py --config cfg3 --devices $NG --check 2>&1 | tee $OUT/group_cfg3_$NG.log | cut -c1-400
fi
if [ "$NG" = "8" ]; then
  echo "== cfg5 plan-per-GPU"; timeout 600 python tools/bench_configs.py --configs cfg5mg --gpus 8 --coils 32 2>&1 | tee $OUT/cfg5_8gpu.log | cut -c1-400
  timeout 600 python tools/bench_configs.py --configs cfg5mg --gpus 1 --coils 32 2>&1 | tee $OUT/cfg5_1gpu.log | cut -c1-400
  echo "== pytest multi"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tee $OUT/pytest_multi.log | tail -5
fi
ls -la $OUT
