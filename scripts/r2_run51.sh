#!/bin/bash
# GPU batch 51: A/B of the bucketed path (sweep points 64 K / 1 M / 100 M groups, sum/min/max/count): the library at 52f74f7 (before
# nullable values rode in the row-number words), the current build, the current build with that handling compiled out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for L in libpa_ab_old.so libpa_b200.so libpa_ab_nonull.so; do
  for G in 65536 1048576 100000000; do
    echo -n "$L G=$G: "
    PA_B200_LIB=$PWD/pandasarrow_b200/lib/$L timeout 200 python scripts/prof_case.py --rows 1000000000 --groups $G --hint $G --aggs sum,min,max,count --iters 4 2>&1 | grep "iter 3" | cut -c1-130
  done
done
