for cfg in "16 1 0" "16 16 24" "16 16 8" "16 16 64" "16 8 24" "16 4 24"; do
  set -- $cfg
  make -C mc_water_ls_mw_b200/csrc -B EXTRA="-DMWGPU_MC_BLOCKS=$1 -DMWGPU_MC_WARPS=$2 -DMWGPU_PACE_SPINS=$3" > /dev/null 2>&1
  echo "warps/SM $1, warps/CTA $2, spins $3: $(grep -A3 k_mc_runILi2 mc_water_ls_mw_b200/csrc/build.log | grep -o 'Used [0-9]* registers') $(timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | grep -o '"value": [0-9.e+]*' | head -1)"
done
