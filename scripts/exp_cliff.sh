make -C mc_water_ls_mw_b200/csrc -B EXTRA="-DMW_EXPERIMENT_NO_ETA" > /dev/null 2>&1
echo "no eta_bin: $(timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | grep -o '"value": [0-9.e+]*' | head -1)"
make -C mc_water_ls_mw_b200/csrc -B > /dev/null 2>&1
echo "baseline: $(timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 | grep -o '"value": [0-9.e+]*' | head -1)"
