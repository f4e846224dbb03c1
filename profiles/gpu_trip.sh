#!/bin/bash
# One GPU trip of the tile-kernel work: parity suite, device-resident bench line, one ncu capture with source counters.
#   gpurun --timeout 900 -- 'bash profiles/gpu_trip.sh <tag> [notests] [noncu]'
tag=$1
if [ "$2" != "notests" ]; then python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${tag}_tests.log; tail -3 gpurun_out/${tag}_tests.log; fi
python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
python -c "import json,sys; d=json.load(open('gpurun_out/${tag}_bench.json')); print('ms_per_step', d['ms_per_step'], 'frac', d['roofline']['frac'], d['config']['timing'])"
if [ "$3" != "noncu" ]; then
ncu --set full --clock-control none --import-source on -k regex:ie_resolve_tile -s 3 -c 1 -o gpurun_out/${tag}_tile -f python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_ncu.log 2>&1
tail -2 gpurun_out/${tag}_ncu.log
fi
