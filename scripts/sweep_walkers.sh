make -C mc_water_ls_mw_b200/csrc -B EXTRA=-DMWGPU_MC_BLOCKS=16 > /dev/null 2>&1
for w in 148 296 592 1184 1776 2368; do
  echo "walkers $w ($((w/148)) per SM): $(timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --walkers $w 2>&1 | tail -1 | grep -o '"value": [0-9.e+]*' | head -1)"
done
