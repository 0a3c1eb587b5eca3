#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 3 python tools/sanitize_small.py > $OUT/sanitize_memcheck.log 2>&1; echo "memcheck rc=$?"; grep -E " ok$|ERROR SUMMARY|Invalid|Error" $OUT/sanitize_memcheck.log | head -30
