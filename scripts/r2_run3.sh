#!/bin/bash
# GPU batch 3: tests, hash-kernel A/B/C, ncu of the hash kernel, bench with sweep
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest3.log 2>&1
tail -3 gpurun_out/r2_pytest3.log
for v in "" _vB _vC; do
  echo "=== variant '$v' ===" >> gpurun_out/r2_abc.log
  PA_B200_LIB=$PWD/pandasarrow_b200/lib/libpa_b200$v.so python scripts/prof_case.py --rows 500000000 --groups 1000 --hashed --iters 4 >> gpurun_out/r2_abc.log 2>&1
  PA_B200_LIB=$PWD/pandasarrow_b200/lib/libpa_b200$v.so python scripts/prof_case.py --rows 500000000 --groups 500 --hashed --iters 3 >> gpurun_out/r2_abc.log 2>&1
done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err
ncu --set full --clock-control none --import-source on -k regex:k_lowcard_scan -c 2 -o gpurun_out/r2_lc_hash python scripts/prof_case.py --rows 200000000 --groups 1000 --hashed --iters 1 > gpurun_out/r2_ncu_hash.log 2>&1
ncu -i gpurun_out/r2_lc_hash.ncu-rep --page raw --csv > gpurun_out/r2_lc_hash_raw.csv 2>/dev/null
