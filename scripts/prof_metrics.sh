#!/bin/bash
# usage: scripts/prof_metrics.sh <tag> <example> <walkers/SM>  -- light ncu pass (issue / instruction-cache counters) of one k_mc_run launch
tag=$1; ex=${2:-ice1_sample}; k=${3:-14}
cd /root/repo
M=gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed.sum.per_cycle_active,sm__icc_requests.sum,sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,sm__warps_active.avg.per_cycle_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio,launch__registers_per_thread
timeout 3000 gpurun --timeout 600 -- "ncu --metrics $M --clock-control none -k regex:k_mc_run -s 8 -c 1 --csv --log-file gpurun_out/met_$tag.csv python scripts/occ_sweep.py $ex $k > gpurun_out/met_$tag.log 2>&1; tail -2 gpurun_out/met_$tag.log" 2>&1 | grep -v "^\[gpurun\] sending\|merged"
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/met_$tag.csv')) if len(r)>10]
hdr=rows[0]
for r in rows[1:]:
    d=dict(zip(hdr,r)); print('%-90s %s %s'%(d['Metric Name'],d['Metric Value'],d['Metric Unit']))
PY
