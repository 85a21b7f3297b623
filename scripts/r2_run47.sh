#!/bin/bash
# GPU batch 47: staged H2D ingest, second sweep: workers x chunk x buffers, every setting twice (the box's host is noisy)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
for rep in 1 2; do
for cfg in "8 8 2" "12 8 2" "12 4 2" "12 4 3" "12 8 3" "16 4 3" "8 4 3" "12 2 4" "12 4 4"; do
  set -- $cfg
  PA_H2D_THREADS=$1 PA_H2D_CHUNK_MB=$2 PA_H2D_BUFS=$3 timeout 120 python scripts/prof_h2d.py --iters 5 2>&1 | tail -1
done; done
