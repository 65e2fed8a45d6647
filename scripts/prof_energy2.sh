#!/bin/bash
# usage: scripts/prof_energy2.sh <tag> -- ncu --set full capture of the batched energy kernel on 65536 thermalised units
tag=$1; cd /root/repo
K=${K:-_ZN2mw2v215k_model_energy3ILi48EEEvNS_11DeviceStateEPd}
bash scripts/gpurun_retry.sh --timeout 900 -- "ncu --set full --clock-control none --import-source on -k regex:k_model_energy3 -s 6 -c 1 -f -o gpurun_out/en2_$tag python scripts/energy_bench.py 8 > gpurun_out/en2_ncu.log 2>&1; tail -3 gpurun_out/en2_ncu.log" 2>&1 | grep -v "^\[gpurun\] sending\|merged" | tail -4
ncu -i gpurun_out/en2_$tag.ncu-rep --page source --csv > gpurun_out/en2_src_$tag.csv 2>/dev/null
ncu -i gpurun_out/en2_$tag.ncu-rep --page raw --csv > gpurun_out/en2_raw_$tag.csv 2>/dev/null
(cd /tmp && cuobjdump -xelf all /root/repo/mc_water_ls_mw_b200/libmwgpu.so >/dev/null 2>&1 && nvdisasm -g -c /tmp/mwgpu.sm_100a.cubin > /tmp/dis_en_$tag.txt 2>/dev/null)
python scripts/ncu_by_line.py gpurun_out/en2_src_$tag.csv /tmp/dis_en_$tag.txt $K 45
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/en2_raw_$tag.csv')))
hdr=rows[0]; r=rows[2]
for h,u,v in zip(hdr,rows[1],r):
    if any(k in h for k in ['issue_stalled','warps_active.avg.per_cycle_active','gpu__time_duration.sum','sm__inst_executed.sum.per_cycle_active','pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread ','launch__occupancy_limit','smsp__inst_executed.sum ','dram__bytes_read.sum ','dram__bytes_write.sum ','dram__throughput.avg.pct']) and 'pcsamp' not in h:
        print(h.replace('smsp__average_warps_issue_stalled_','stall_').replace('_per_issue_active.ratio',''),u,v)
PY
