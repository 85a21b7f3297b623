#!/bin/bash
# GPU batch 52: final build of the round (NULLABLE as a template parameter of the bucketed kernels, fp32 bounds pre-check): smoke(), the whole GPU suite, the default bench line, launch list of the bench step
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1100 python -m pytest tests -m gpu -q -x > gpurun_out/r2_pytest52.log 2>&1
echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest52.log
timeout 400 python bench.py > gpurun_out/r2_bench52.json 2> gpurun_out/r2_bench52.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/r2_bench52.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-sweep --no-extras > /dev/null 2>&1
ls -la gpurun_out | grep -E "bench52|launches"
