#!/bin/bash
# GPU batch 19: evidence for profiles/ — bench line, launch list of the bench command, ncu --set full of the bucket
# aggregation (ranged and whole-bucket), racecheck of the lowcard hash-mode insert path
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
(time python bench.py) > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err; tail -2 gpurun_out/r2_bench_c.err
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-sweep --no-extras > gpurun_out/r2_bench_short.json 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-sweep --no-extras > gpurun_out/r2_bench_ncu.log 2>&1
for G in 65536 1048576; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_bucket_agg -s 1 -c 1 -o gpurun_out/r2_bkagg_$G -f python scripts/prof_bucketed.py --rows 1000000000 --iters 2 --groups $G > gpurun_out/r2_bkagg_$G.log 2>&1
  ncu -i gpurun_out/r2_bkagg_$G.ncu-rep --page raw --csv > gpurun_out/r2_bkagg_${G}_raw.csv 2>/dev/null
  python scripts/ncu_summary.py gpurun_out/r2_bkagg_${G}_raw.csv > gpurun_out/r2_bkagg_$G.md
  rm -f gpurun_out/r2_bkagg_$G.ncu-rep
done
# racecheck: hash mode (scattered keys), small input so that the insert path is a large share of the run
timeout 900 compute-sanitizer --tool racecheck --racecheck-report all python scripts/prof_case.py --rows 300000 --groups 1000 --hashed --iters 1 > gpurun_out/r2_racecheck_lowcard_hash.log 2>&1
tail -5 gpurun_out/r2_racecheck_lowcard_hash.log
