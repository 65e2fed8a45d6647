#!/bin/bash
# usage: scripts/sanitize.sh <tag>  -- compute-sanitizer racecheck + memcheck of the walker kernels (both generations),
# the service kernels and the bin reduction on a B200 (gpurun); summaries are copied to profiles/<tag>_sanitize_*.txt
tag=${1:-r2}
cd /root/repo
CS=/usr/local/cuda/bin/compute-sanitizer
run() {  # tool deck walkers cycles kernel
  echo "$CS --tool $1 --print-limit 20 python scripts/sanitize_run.py $2 $3 $4 $5 > gpurun_out/san_${1}_${2}_k$5.txt 2>&1;"
}
CMD="$(run racecheck ice1_sample 16 6 2) $(run racecheck ice1_gen_weights 16 6 2) $(run racecheck single_box 16 6 2) $(run racecheck ice1_sample 16 6 1) \
$(run memcheck ice1_sample 32 10 2) $(run memcheck ice1_gen_weights 32 10 2) $(run memcheck single_box 32 10 2) $(run memcheck ice1_sample 32 10 1) \
tail -n 4 gpurun_out/san_*.txt"
timeout 3000 gpurun --timeout 1500 -- "$CMD" 2>&1 | grep -v "^\[gpurun\] sending\|merged" | tail -60
for f in gpurun_out/san_*_k*.txt; do
  b=$(basename $f .txt)
  { echo "# $b: $(grep -c 'Race reported\|Invalid\|Error:' $f) findings"; grep -E "^ok |RACECHECK SUMMARY|ERROR SUMMARY|Race reported|Invalid __" $f | sort | uniq -c | head -30; } > profiles/${tag}_${b}.txt
done
