#!/bin/bash
# GPU batch 43: resample min / max folded as doubles (DSETP + SEL): parity (resample suite incl. the new special-values test,
# config 4 at 100 M / 1 B ticks, golden, facade), timing, then ncu --set full of k_resample_scan (OHLC + sum; sum/mean/count)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_resample_gpu.py tests/test_zz_golden_gpu.py tests/test_facade_gpu.py -m gpu -q -x > gpurun_out/r2_pytest43.log 2>&1
echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest43.log
timeout 600 python -m pytest tests/test_parity_large_gpu.py tests/test_fullsize_gpu.py -m gpu -q -x -k "config4 or resample" 2>&1 | tail -3
timeout 300 python scripts/prof_resample.py --rows 1000000000 --iters 3 2>&1 | grep "iter [12]" | cut -c1-200
timeout 300 python scripts/prof_resample.py --rows 1000000000 --iters 3 --aggs sum,mean,count 2>&1 | grep "iter 2" | cut -c1-200
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resample_scan -c 1 -o gpurun_out/r2_rs_ohlc python scripts/prof_resample.py --rows 1000000000 --iters 1 > gpurun_out/r2_ncu_rs_ohlc.log 2>&1
ncu -i gpurun_out/r2_rs_ohlc.ncu-rep --page raw --csv > gpurun_out/r2_rs_ohlc_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_rs_ohlc.ncu-rep --page source --csv > gpurun_out/r2_rs_ohlc_src.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_resample_scan -c 1 -o gpurun_out/r2_rs_narrow python scripts/prof_resample.py --rows 1000000000 --iters 1 --aggs sum,mean,count > gpurun_out/r2_ncu_rs_narrow.log 2>&1
ncu -i gpurun_out/r2_rs_narrow.ncu-rep --page raw --csv > gpurun_out/r2_rs_narrow_raw.csv 2>/dev/null
rm -f gpurun_out/r2_rs_ohlc.ncu-rep gpurun_out/r2_rs_narrow.ncu-rep
ls -la gpurun_out | grep r2_rs
