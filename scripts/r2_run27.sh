#!/bin/bash
# GPU batch 27: few SCATTERED keys (what hashed utf8 keys with a handful of values look like)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for G in 2 8 32 128; do
echo "== $G random 64-bit keys"
timeout 300 python scripts/prof_case.py --rows 1000000000 --groups $G --hashed --iters 3 2>&1 | grep "iter [2]" | cut -c1-120
timeout 300 python scripts/prof_case.py --rows 1000000000 --groups $G --hashed --iters 3 --path global 2>&1 | grep "iter [2]" | cut -c1-120
done
