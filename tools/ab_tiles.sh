#!/bin/bash
# same-box A/B of two builds of the library (build/libub200_old.so vs build/libub200_new.so) on the splat path:
# 1 M Gaussians at 1297x840 (tools/perf_kernels.py splat), alternating builds
P=uncertainty_nerf_gs_b200
for round in 1 2 3; do
  for v in old new; do
    cp $P/build/libub200_$v.so $P/libub200.so
    echo -n "$v: "
    timeout 120 python tools/perf_kernels.py splat 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k: round(v, 4) for k, v in d.items() if k.startswith('ms_')})"
  done
done
cp $P/build/libub200_new.so $P/libub200.so
