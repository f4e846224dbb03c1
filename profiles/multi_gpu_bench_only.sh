#!/bin/bash
# One 8-GPU box: the bench line at N = 2, 4, 8 (one process per GPU under torchrun, plus the single-process
# ie_resolve_batch_multi leg on rank 0).  The copy probes of profiles/multi_gpu_trip.sh are not repeated.
#   gpurun --gpus 8 --timeout 900 -- 'bash profiles/multi_gpu_bench_only.sh r02f'
tag=$1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
for n in 2 4 8; do
  run $n 2970$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_${n}gpu.json 2> gpurun_out/${tag}_bench_${n}gpu.err
  python -c "import json; d=json.load(open('gpurun_out/${tag}_bench_${n}gpu.json')); print($n, 'value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'single_process', (d['e2e'].get('single_process') or {}).get('value'))"
done
