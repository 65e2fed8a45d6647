#!/bin/bash
# usage: scripts/prof_final.sh <tag>  -- the evidence set of a build: plain bench line, ncu launch list of the same
# command, one --set full capture of the dominant kernel (k_mc_run, one 250-cycle step launch) and one of
# k_model_energy_all; summaries are written under profiles/ (B200_PROFILING.md recipe; numbers under ncu are never bench values)
set -e
tag=$1
cd /root/repo
K2=_ZN2mw2v29k_mc_run2ILi2ELi48ELi14ELi1ELi1EEEvNS_11DeviceStateENS_8McParamsEii
bash scripts/gpurun_retry.sh --timeout 2400 -- "python bench.py --steps 10 --warmup 3 > gpurun_out/final_bench_$tag.json 2> gpurun_out/final_bench_$tag.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/final_launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/final_ncu1.log 2>&1; \
ncu --set full --clock-control none --import-source on -k regex:k_mc_run2 -s 6 -c 1 -f -o gpurun_out/final_mc_$tag python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/final_ncu2.log 2>&1; \
ncu --set full --clock-control none --import-source on -k regex:k_model_energy4 -s 2 -c 1 -f -o gpurun_out/final_en_$tag python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/final_ncu3.log 2>&1; \
tail -c 400 gpurun_out/final_bench_$tag.json" 2>&1 | tail -5
for k in mc en; do
  ncu -i gpurun_out/final_${k}_$tag.ncu-rep --page raw --csv > profiles/${tag}_${k}_raw.csv 2>/dev/null
  ncu -i gpurun_out/final_${k}_$tag.ncu-rep --page source --csv > gpurun_out/final_${k}_src_$tag.csv 2>/dev/null
done
(cd /tmp && cuobjdump -xelf all /root/repo/mc_water_ls_mw_b200/libmwgpu.so >/dev/null 2>&1 && nvdisasm -g -c /tmp/mwgpu.sm_100a.cubin > /tmp/dis_$tag.txt 2>/dev/null)
# one launch = 250 cycles x 48 moves x 4096 walkers
python scripts/ncu_by_line.py gpurun_out/final_mc_src_$tag.csv /tmp/dis_$tag.txt $K2 70 > profiles/${tag}_k_mc_run_by_line.txt 2>&1 || true
python scripts/ncu_hotset.py gpurun_out/final_mc_src_$tag.csv /tmp/dis_$tag.txt $K2 49152000 50 > profiles/${tag}_k_mc_run_hotset.txt 2>&1 || true
cp gpurun_out/final_bench_$tag.json profiles/${tag}_bench.json
cp gpurun_out/final_launches_$tag.csv profiles/${tag}_launches.csv
python - <<PY
import csv, json
def pick(path, keys):
    rows=list(csv.reader(open(path))); hdr=rows[0]; r=rows[-1]
    return {h:(r[i], rows[1][i]) for i,h in enumerate(hdr) if h in keys}
keys=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','smsp__inst_executed.sum','sm__inst_executed.sum.per_cycle_active',
      'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','sm__icc_request_hit_rate.pct','gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed',
      'launch__registers_per_thread','sm__warps_active.avg.per_cycle_active','dram__bytes.sum.per_second']
out={}
for k in ('mc','en'):
    out[k]=pick('profiles/${tag}_%s_raw.csv'%k, keys)
    print(k, json.dumps(out[k], indent=0))
def tobytes(v,u):
    v=float(v); return v*{'byte':1,'Kbyte':1e3,'Mbyte':1e6,'Gbyte':1e9}[u]
t={}
for k,name in (('mc','k_mc_run'),('en','k_model_energy')):
    d=out[k]
    t[name]={'dram_bytes_read':tobytes(*d['dram__bytes_read.sum']),'dram_bytes_write':tobytes(*d['dram__bytes_write.sum']),
             'duration_ms_under_ncu':float(d['gpu__time_duration.sum'][0])*{'ms':1,'us':1e-3,'ns':1e-6,'s':1e3}[d['gpu__time_duration.sum'][1]]}
    t[name]['dram_bytes']=t[name]['dram_bytes_read']+t[name]['dram_bytes_write']
    t[name]['warp_instructions']=float(d['smsp__inst_executed.sum'][0])
    t[name]['ipc_per_sm_under_ncu']=float(d['sm__inst_executed.sum.per_cycle_active'][0])/148.0
t['source']='ncu --set full --clock-control none, one launch of the default bench.py command (k_mc_run = v2::k_mc_run2<2,48>, 4096 walkers, 250 cycles per launch; k_model_energy = v2::k_model_energy4<48>, 65536 evaluations per launch); profiles/${tag}_*_raw.csv'
import sys; sys.path.insert(0,'/root/repo')
import bench
t['csrc_sha256']=bench.csrc_sha256()
json.dump(t, open('profiles/traffic.json','w'), indent=1)
print(json.dumps(t, indent=1))
PY
