#!/bin/bash
# GPU batch 4: tests, hash-kernel A/B/C, bench with sweep, ncu of the hash kernel
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
nvidia-smi --query-gpu=name,memory.used,memory.total --format=csv > gpurun_out/r2_smi4.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest4.log 2>&1
tail -3 gpurun_out/r2_pytest4.log
rm -f gpurun_out/r2_abc.log
for v in "" _vB _vC; do
  echo "=== variant '$v' ===" >> gpurun_out/r2_abc.log
  PA_B200_LIB=$PWD/pandasarrow_b200/lib/libpa_b200$v.so timeout 300 python scripts/prof_case.py --rows 500000000 --groups 1000 --hashed --iters 4 >> gpurun_out/r2_abc.log 2>&1
  PA_B200_LIB=$PWD/pandasarrow_b200/lib/libpa_b200$v.so timeout 300 python scripts/prof_case.py --rows 500000000 --groups 500 --hashed --iters 3 >> gpurun_out/r2_abc.log 2>&1
done
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_lowcard_scan -c 2 -o gpurun_out/r2_lc_hash python scripts/prof_case.py --rows 200000000 --groups 1000 --hashed --iters 1 > gpurun_out/r2_ncu_hash.log 2>&1
ncu -i gpurun_out/r2_lc_hash.ncu-rep --page raw --csv > gpurun_out/r2_lc_hash_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_lc_hash.ncu-rep --page source --csv > gpurun_out/r2_lc_hash_src.csv 2>/dev/null
