#!/usr/bin/env python
"""Extract one function's SASS from `cuobjdump -sass`: python scripts/sass_fn.py <lib.so> <mangled-substring>"""
import subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
keep = False
for line in out.splitlines():
    if "Function :" in line:
        keep = sys.argv[2] in line
    if keep and ("/*0" in line or "Function" in line or "/*1" in line or "/*2" in line) and not line.strip().startswith("/* 0x"):
        print(line.split("/* 0x")[0].rstrip())
