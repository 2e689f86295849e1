#!/bin/bash
mkdir -p gpurun_out
L=$PWD/tensoralloy_b200/csrc/libtab200_td8k.so
for cfg in "16 4" "8 4" "8 2"; do set -- $cfg
  echo "8 KB stages, tile $1 split $2"; TAB200_LIB=$L TAB_TD_TILE=$1 TAB_TD_SPLIT=$2 timeout 300 python tools/td_heads_bench.py 65536 2>&1 | tail -2 | cut -c1-175
done | tee gpurun_out/r02zz_td_heads_8k.log
