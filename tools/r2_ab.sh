#!/bin/bash
# A/B of library variants on a few shapes/configs: tools/r2_ab.sh "<variants>" "<shapes>" "<cfgs>"
for v in $1; do
  if [ "$v" = base ]; then unset PQ_LIB_PATH; else export PQ_LIB_PATH=$PWD/protoquant_b200/libpq_$v.so; fi
  echo "== variant $v"
  PQ_SKIP_LIBS=1 PQ_SHAPES="$2" PQ_CFGS="$3" python tools/sweep_r2.py hot 2>&1 | grep -v "^mode"
done
