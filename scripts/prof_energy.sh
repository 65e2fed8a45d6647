#!/bin/bash
# usage: scripts/prof_energy.sh <tag>  -- full ncu capture of one k_model_energy_all launch + per-line summary (development aid)
tag=$1
cd /root/repo
timeout 3000 gpurun --timeout 900 -- "ncu --set full --clock-control none --import-source on -k regex:k_model_energy_all -s 2 -c 1 -f -o gpurun_out/en_$tag python bench.py --steps 1 --warmup 3 --no-cpu --cycles 10 > gpurun_out/en_ncu.log 2>&1; tail -1 gpurun_out/en_ncu.log | cut -c1-100" 2>&1 | grep -v "^\[gpurun\] sending\|merged"
ncu -i gpurun_out/en_$tag.ncu-rep --page source --csv > gpurun_out/en_src_$tag.csv 2>/dev/null
ncu -i gpurun_out/en_$tag.ncu-rep --page raw --csv > gpurun_out/en_raw_$tag.csv 2>/dev/null
(cd /tmp && cuobjdump -xelf all /root/repo/mc_water_ls_mw_b200/libmwgpu.so >/dev/null 2>&1 && nvdisasm -g -c /tmp/mwgpu.sm_100a.cubin > /tmp/dis_en_$tag.txt 2>/dev/null)
python scripts/ncu_by_line.py gpurun_out/en_src_$tag.csv /tmp/dis_en_$tag.txt _Z18k_model_energy_allN2mw11DeviceStateEPd 40
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/en_raw_$tag.csv'))); hdr=rows[0]; r=rows[-1]
for h,u,v in zip(hdr,rows[1],r):
    if h in ('gpu__time_duration.sum','smsp__inst_executed.sum','sm__inst_executed.sum.per_cycle_active','sm__warps_active.avg.per_cycle_active','launch__registers_per_thread','launch__occupancy_limit_shared_mem','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__thread_inst_executed_per_inst_executed.ratio') or ('issue_stalled' in h and 'per_issue_active' in h and float(v)>0.1): print(h.replace('smsp__average_warps_issue_stalled_','stall_'),u,v)
PY
