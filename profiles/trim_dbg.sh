#!/bin/bash
# timing decomposition of the trimmed int8-slice launches (RDB_OZAKI_DBG: 1 = no TMA loads, 2 = no epilogue; results are garbage)
for tp in ${TPS:-0 1}; do for dbg in 0 1 2 3; do
  echo "== RDB_OZAKI_TPAIR=$tp RDB_OZAKI_DBG=$dbg"
  RDB_OZAKI_TPAIR=$tp RDB_OZAKI_DBG=$dbg CASES=-1:0,0:-1,-1:-1 REPS=10 python profiles/trim_probe.py 2>&1 | head -3 | cut -c1-60
done; done
