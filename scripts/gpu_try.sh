#!/bin/bash
# usage: scripts/gpu_try.sh <tag> [notest]  -- parity tests + short bench of the current build on a B200 (development aid)
tag=$1
cd /root/repo
T='python -m pytest tests -m gpu -x -q 2>&1 | tail -3;'
[ "$2" = "notest" ] && T=''
bash scripts/gpurun_retry.sh --timeout 900 -- "$T python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_$tag.json 2>&1; python -c \"
import json;d=json.loads(open('gpurun_out/bench_$tag.json').read().strip().splitlines()[-1]);print('moves/s %.4g  k_ms %.3f  e2e %.4g  evals/s %.4g'%(d['value'],d['roofline']['kernel_ms'],d['e2e']['value'],d['energy_evals']['value']))\" || tail -5 gpurun_out/bench_$tag.json" 2>&1 | grep -v "^\[gpurun\] sending\|merged"
