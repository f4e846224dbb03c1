#!/bin/bash
# One 8-GPU box: the N-rank copy probe, the bench line at N = 2, 4, 8 (one process per GPU under torchrun, plus the
# single-process ie_resolve_batch_multi leg on rank 0), the single-process copy probe.
#   gpurun --gpus 8 --timeout 900 -- 'bash profiles/multi_gpu_trip.sh r02'
tag=$1
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
nvidia-smi topo -m > gpurun_out/${tag}_topo.txt 2>&1
lscpu | head -25 > gpurun_out/${tag}_lscpu.txt 2>&1
for n in 1 2 4 8; do run $n 2960$n profiles/copy_probe.py 2>/dev/null | tail -1 >> gpurun_out/${tag}_copy_probe.jsonl; done
python profiles/copy_probe.py --single 8 2>/dev/null | tail -1 >> gpurun_out/${tag}_copy_probe.jsonl
python profiles/copy_probe.py --single 4 2>/dev/null | tail -1 >> gpurun_out/${tag}_copy_probe.jsonl
cat gpurun_out/${tag}_copy_probe.jsonl | cut -c1-420
for n in 2 4 8; do
  run $n 2970$n bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_${n}gpu.json 2> gpurun_out/${tag}_bench_${n}gpu.err
  python -c "import json; d=json.load(open('gpurun_out/${tag}_bench_${n}gpu.json')); print($n, 'value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'single_process', d['e2e'].get('single_process'))"
done
