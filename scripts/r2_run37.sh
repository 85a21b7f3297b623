#!/bin/bash
# GPU batch 37: the whole GPU suite, smoke, the default bench line, the launch list of the bench command
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/r2_pytest_full3.log 2>&1
tail -6 gpurun_out/r2_pytest_full3.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time timeout 900 python bench.py ) > gpurun_out/r2_bench_f.json 2> gpurun_out/r2_bench_f.err
tail -c 600 gpurun_out/r2_bench_f.json; tail -4 gpurun_out/r2_bench_f.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-sweep --no-extras > gpurun_out/r2_bench_ncu.log 2>&1
tail -2 gpurun_out/r2_bench_ncu.log | cut -c1-300
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 ) 2>&1 | tail -6 | cut -c1-600
