#!/bin/bash
# Replicated-forward-sweep measurement on N GPUs (default 2): multi-rank tests (N = 2 only, MG_NO_TESTS skips them),
# in-kernel counters in replicate mode, then bench.py strong scaling in replicate (default) and shard mode.
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ] && [ -z "$MG_NO_TESTS" ]; then
  timeout 600 python -m pytest tests/test_multigpu.py -x -q --durations=5 ${MG_TESTS_K:+-k "$MG_TESTS_K"} > gpurun_out/s3_mg_tests_$N.txt 2>&1; echo "tests rc=$?"; tail -12 gpurun_out/s3_mg_tests_$N.txt
fi
for direct in "" 1; do  # forwarder warps (default), then every producing warp storing to all ranks itself
  KROTOV_RF_DIRECT=$direct MG_MODE=replicate timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 tools/gpu_mg_prof.py > gpurun_out/s3_mg_prof_rf${direct:+_direct}_$N.txt 2>&1; echo "prof direct=$direct rc=$?"
  grep -E "exchange=|wait_pulse|comm_gather|backward|forward " gpurun_out/s3_mg_prof_rf${direct:+_direct}_$N.txt
done
run() {  # name, bench args...
  name=$1; shift
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 \
      bench.py --gpus $N --steps 10 --warmup 3 "$@" > gpurun_out/s3_mg_${name}_$N.json 2> gpurun_out/s3_mg_${name}_$N.err
  echo "$name rc=$? $(python -c "import json,sys; d=json.load(open('gpurun_out/s3_mg_${name}_$N.json')); print(d['ms_per_step'], d['config']['grid'], d['e2e']['iterations_per_s'], d['config'].get('parallelism'), d['config']['J_T_last'])" 2>&1 | tail -1)"
}
for x in ${MG_BENCH:-auto shard}; do run $x --multi-gpu $x; done
