#!/bin/bash
# GPU batch 46: staged (pageable) H2D ingest: worker-thread and staging-chunk sweep (one process per setting)
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
nproc; grep -m1 "model name" /proc/cpuinfo
for T in 4 8 12 16; do PA_H2D_THREADS=$T timeout 120 python scripts/prof_h2d.py 2>&1 | tail -1; done
for C in 1 2 4 16 32; do PA_H2D_THREADS=8 PA_H2D_CHUNK_MB=$C timeout 120 python scripts/prof_h2d.py 2>&1 | tail -1; done
for C in 2 32; do PA_H2D_THREADS=16 PA_H2D_CHUNK_MB=$C timeout 120 python scripts/prof_h2d.py 2>&1 | tail -1; done
