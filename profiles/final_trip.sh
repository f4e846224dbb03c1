#!/bin/bash
# End-of-round measurement trip on one B200: parity suite, the contract line (with e2e and the CPU baseline), the reference
# arm, the other workloads, the launch list of the bench command and one full ncu capture of the hot kernel.
#   gpurun --timeout 1500 -- 'bash profiles/final_trip.sh r02'
tag=$1
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/${tag}_final_tests.log; tail -2 gpurun_out/${tag}_final_tests.log
python bench.py --steps 20 --warmup 5 > gpurun_out/${tag}_bench_final_1gpu.json 2> gpurun_out/${tag}_bench_final_1gpu.err
python -c "import json; d=json.load(open('gpurun_out/${tag}_bench_final_1gpu.json')); print('ms_per_step', d['ms_per_step'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'cpu', d['cpu_baseline']['value'])"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference_arm.json 2>/dev/null; cut -c1-300 gpurun_out/${tag}_bench_reference_arm.json
for w in c3 c5 escape c1c2; do python bench.py --workload $w > gpurun_out/${tag}_bench_${w}.json 2>/dev/null; python -c "import json; d=json.load(open('gpurun_out/${tag}_bench_${w}.json')); print('$w', d.get('value'), d.get('ms_per_step'), d['roofline'].get('frac') if d.get('roofline') else None)"; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_final_launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ie_resolve_fused -s 3 -c 1 -o gpurun_out/${tag}_final_fused -f python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_final_ncu.log 2>&1
tail -2 gpurun_out/${tag}_final_ncu.log
