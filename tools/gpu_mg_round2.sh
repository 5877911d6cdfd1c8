#!/bin/bash
# Round-2 multi-GPU measurement on N GPUs (default 2): multi-rank tests (N = 2 only), in-kernel counters of the three
# exchange protocols, then bench.py strong scaling per protocol.
N=${1:-2}
mkdir -p gpurun_out
if [ "$N" = "2" ] && [ -z "$MG_NO_TESTS" ]; then
  timeout 600 python -m pytest tests/test_multigpu.py -x -q > gpurun_out/r2_mg_tests_$N.txt 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_mg_tests_$N.txt
fi
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 tools/gpu_mg_prof.py > gpurun_out/r2_mg_prof_$N.txt 2>&1; echo "prof rc=$?"
grep -E "exchange=|wait_pulse|comm_gather|backward|forward " gpurun_out/r2_mg_prof_$N.txt
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 \
      bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_mg_${name}_$N.json 2> gpurun_out/r2_mg_${name}_$N.err
  echo "$name rc=$? $(python -c "import json,sys; d=json.load(open('gpurun_out/r2_mg_${name}_$N.json')); print(d['ms_per_step'], d['config']['grid'], d['e2e']['iterations_per_s'], d['config'].get('exchange'))" 2>&1 | tail -1)"
}
for x in ${MG_BENCH:-hier hierst onehop}; do run $x KROTOV_XCHG=$x; done
