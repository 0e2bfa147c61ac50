#!/bin/bash
# 8-GPU box: multi-GPU tests + the bench line at N = 2, 4, 8 (weak C1 value + strong C5 sub-record).  bash tools/run_scaling.sh <tag>
TAG=${1:-r3}
python -m pytest tests/test_gpu_multi.py tests/test_cpp_mirror.py -m gpu -q -s 2>&1 | tail -12 > gpurun_out/${TAG}_multi_tests.log
cat gpurun_out/${TAG}_multi_tests.log | tail -6
for N in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err
  python -c "import json; d=json.loads(open('gpurun_out/${TAG}_bench_n$N.json').read().strip().splitlines()[-1]); s=d['strong_scaling']; print($N, round(d['value']), round(d['ms_per_step'],1), d['reduce']['ms_per_step_rank0'], 'strong', round(s['value']), round(s['ms_per_step']), s['reduce_ms_rank0'])" 2>&1 | tail -1
done
