for n in 14 16 20 24; do
  make -C mc_water_ls_mw_b200/csrc -B EXTRA=-DMWGPU_MC_BLOCKS=$n > /dev/null 2>&1
  echo "blocks/SM cap $n: $(grep -A3 k_mc_runILi2 mc_water_ls_mw_b200/csrc/build.log | grep Used)"
  timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | grep -o '"value": [0-9.e+]*' | head -1
done
