#!/bin/bash
# GPU batch 26: hash-mode variants (next row group requested after / before the probes)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
cp pandasarrow_b200/lib/libpa_b200.so /tmp/orig.so
for V in early0 early1; do
  cp pandasarrow_b200/lib/variant_$V.so pandasarrow_b200/lib/libpa_b200.so
  echo "== $V"
  timeout 300 python scripts/prof_case.py --rows 1000000000 --groups 1000 --hashed --iters 4 2>&1 | grep "iter [23]" | cut -c1-120
done
cp /tmp/orig.so pandasarrow_b200/lib/libpa_b200.so
