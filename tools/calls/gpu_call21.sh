#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/td_heads_bench.py 65536 > gpurun_out/r02r_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_td_heads' -s 3 -c 1 -f -o gpurun_out/r02r_td_heads python tools/td_heads_bench.py 65536 > gpurun_out/r02r_ncu.log 2>&1
ls -la gpurun_out/r02r*
