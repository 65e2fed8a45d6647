for cfg in "16 1" "16 4" "16 8" "16 16" "12 12"; do
  set -- $cfg
  make -C mc_water_ls_mw_b200/csrc -B EXTRA="-DMWGPU_MC_BLOCKS=$1 -DMWGPU_MC_WARPS=$2" > /dev/null 2>&1
  echo "warps/SM $1, warps/CTA $2: $(grep -A3 k_mc_runILi2 mc_water_ls_mw_b200/csrc/build.log | grep -o 'Used [0-9]* registers') $(timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | grep -o '"value": [0-9.e+]*' | head -1)"
done
make -C mc_water_ls_mw_b200/csrc -B EXTRA="-DMWGPU_MC_BLOCKS=16 -DMWGPU_MC_WARPS=4" > /dev/null 2>&1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
