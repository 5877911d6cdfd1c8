#!/bin/bash
# Ad-hoc: multi-GPU exchange variants on N GPUs (default 2): tests, then bench.py strong scaling with
# the one-hop cross-rank sum (accumulator word stride 1 and 16) and the mailbox protocol.
N=${1:-2}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_multigpu.py -x -q > gpurun_out/mg_tests_$N.txt 2>&1; echo "tests rc=$?"
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 \
      bench.py --gpus $N --steps 6 --warmup 3 > gpurun_out/mg_${name}_$N.json 2> gpurun_out/mg_${name}_$N.err
  echo "$name rc=$? $(python -c "import json,sys; d=json.load(open('gpurun_out/mg_${name}_$N.json')); print(d['ms_per_step'], d['config']['grid'], d['e2e']['iterations_per_s'])" 2>&1 | tail -1)"
}
run xacc1 KROTOV_XACC_STRIDE=1
run xacc16 KROTOV_XACC_STRIDE=16
run mbox KROTOV_NO_XACC=1
