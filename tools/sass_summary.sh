#!/bin/bash
# SASS evidence: per kernel of libub200.so, how often the instructions that prove the design show up --
#   UBLKCP (cp.async.bulk: TMA 1-D bulk copies), SYNCS (mbarrier), UTCHMMA / LDTM / UTCBAR (tcgen05 MMA, TMEM load,
#   tcgen05 commit), MUFU.EX2, shared / global atomics, FP64 adds.  Usage: tools/sass_summary.sh > profiles/rN_sass_summary.txt
set -e
LIB="$(dirname "$0")/../uncertainty_nerf_gs_b200/libub200.so"
echo "# cuobjdump -sass $(basename "$LIB") | per-kernel counts of selected mnemonics ($(date -u +%F), $(nvcc --version | tail -1))"
cuobjdump -sass "$LIB" | awk '
  /Function :/ {fn=$3}
  {
    for (i = 1; i <= NF; ++i) {
      t = $i
      if (t ~ /^(UBLKCP|UTCHMMA|UTCMMA|LDTM|STTM|UTCBAR|UTMALDG|SYNCS|ATOMS|ATOMG|RED|DADD|DFMA|DMUL)(\.|$)/) { k = t; sub(/\..*/, "", k); cnt[fn " " k]++ }
      else if (t ~ /^MUFU\.EX2/) cnt[fn " MUFU.EX2"]++
    }
  }
  END { for (x in cnt) print x, cnt[x] }' | sort | c++filt | awk '
  { n = $NF; k = $(NF - 1); $NF = ""; $(NF - 1) = ""; sub(/ +$/, ""); rows[$0] = rows[$0] " " k "=" n }
  END { for (f in rows) print f ":" rows[f] }' | sort
echo "# totals"
cuobjdump -sass "$LIB" | grep -oE "\b(UBLKCP|UTCHMMA|LDTM|UTCBAR|UTMALDG|SYNCS)\b" | sort | uniq -c
