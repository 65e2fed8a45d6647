#!/bin/bash
# usage: scripts/prof.sh <tag>   -- ncu capture of k_mc_run under gpurun, then per-line / hot-set summaries here
set -e
tag=$1
cd /root/repo
bash scripts/gpurun_retry.sh --timeout 1200 -- "python bench.py --steps 3 --warmup 3 --no-cpu --cycles 10 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_mc_run -s 8 -c 1 -f -o gpurun_out/prof_$tag python bench.py --steps 3 --warmup 3 --no-cpu --cycles 10 > gpurun_out/ncu.log 2>&1; tail -c 3000 gpurun_out/plain.log | grep -o '\"value\": [0-9.e+]*' | head -1" 2>&1 | tail -4
ncu -i gpurun_out/prof_$tag.ncu-rep --page source --csv > gpurun_out/src_$tag.csv 2>/dev/null
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/raw_$tag.csv 2>/dev/null
(cd /tmp && cuobjdump -xelf all /root/repo/mc_water_ls_mw_b200/libmwgpu.so >/dev/null 2>&1 && nvdisasm -g -c /tmp/mwgpu.sm_100a.cubin > /tmp/dis_$tag.txt 2>/dev/null)
K=${K:-_ZN2mw2v29k_mc_run2ILi2ELi48EEEvNS_11DeviceStateENS_8McParamsEi}
python scripts/ncu_by_line.py gpurun_out/src_$tag.csv /tmp/dis_$tag.txt $K 60 > gpurun_out/byline_$tag.txt 2>&1 || true
python scripts/ncu_hotset.py gpurun_out/src_$tag.csv /tmp/dis_$tag.txt $K 1966080 50 > gpurun_out/hotset_$tag.txt 2>&1 || true
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/raw_$tag.csv')))
hdr=rows[0]; r=rows[2]
for h,u,v in zip(hdr,rows[1],r):
    if any(k in h for k in ['issue_stalled','warps_active.avg.per_cycle_active','gpu__time_duration.sum','sm__inst_executed.sum.per_cycle_active','pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread ','launch__occupancy_limit','smsp__inst_executed.sum ']) and 'pcsamp' not in h:
        print(h.replace('smsp__average_warps_issue_stalled_','stall_').replace('_per_issue_active.ratio',''),u,v)
PY
head -12 gpurun_out/hotset_$tag.txt
