#!/bin/bash
# GPU box: the bench.py line of every config at its FULL spp (BASELINE.json), one GPU.  Output: profiles/<round>/bench_C*.json
# (via gpurun_out/).  Usage: bash tools/run_all_benches.sh <tag>
TAG=${1:-r3}
mkdir -p gpurun_out
python bench.py --workload C1 --steps 3 --warmup 3                    > gpurun_out/${TAG}_bench_C1.json 2> gpurun_out/${TAG}_bench_C1.err
python bench.py --workload C2 --steps 3 --warmup 3 --strong-spp 0     > gpurun_out/${TAG}_bench_C2.json 2> gpurun_out/${TAG}_bench_C2.err
python bench.py --workload C3 --steps 1 --warmup 3 --strong-spp 0     > gpurun_out/${TAG}_bench_C3.json 2> gpurun_out/${TAG}_bench_C3.err
python bench.py --workload C4 --steps 2 --warmup 3 --strong-spp 0     > gpurun_out/${TAG}_bench_C4.json 2> gpurun_out/${TAG}_bench_C4.err
python bench.py --workload C5 --spp 2048 --steps 1 --warmup 3 --strong-spp 0 --no-cpu-baseline > gpurun_out/${TAG}_bench_C5.json 2> gpurun_out/${TAG}_bench_C5.err
python bench.py --impl reference --steps 2 --warmup 0                 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
for f in gpurun_out/${TAG}_bench_*.json; do echo "$f $(python -c "import json,sys; d=json.loads(open('$f').read().strip().splitlines()[-1]); print(round(d['value'],1), d.get('ms_per_step'))" 2>&1)"; done
