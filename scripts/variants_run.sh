#!/bin/bash
# time every variants/libmfgp_*.so on the bench workload
for f in variants/libmfgp_*.so; do
  echo "== $f"; MFGP_LIB_PATH=$PWD/$f python scripts/small_once.py ${1:-16384} 3 | tail -3
done
