#!/bin/bash
cd /root/repo
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_full_r2.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_full_r2.txt
timeout 100 python tools/sanitize_pass.py > gpurun_out/sanitize_plain.txt 2>&1 && timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_pass.py > gpurun_out/memcheck_r2.txt 2>&1; echo "memcheck rc=$?" >> gpurun_out/memcheck_r2.txt
tail -n 8 gpurun_out/pytest_full_r2.txt; tail -n 12 gpurun_out/memcheck_r2.txt
