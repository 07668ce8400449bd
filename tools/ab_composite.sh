#!/bin/bash
# same-box A/B of two builds of the library (build/libub200_old.so vs build/libub200_new.so): the compositing call on one
# full view, alternating builds (box-to-box variation is +-3 %, larger than most tuning steps)
P=uncertainty_nerf_gs_b200
for round in 1 2 3; do
  for v in old new; do
    cp $P/build/libub200_$v.so $P/libub200.so
    echo -n "$v: "
    UB_TUNE_CFGS="7" bash tools/tune_composite.sh 2>&1 | grep NCW
  done
done
cp $P/build/libub200_new.so $P/libub200.so
