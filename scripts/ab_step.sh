#!/usr/bin/env bash
# A/B timing of library variants on one box: scripts/ab_step.sh c3 5 base cta8 ...   ("base" = the in-tree library)
W="$1"; S="$2"; shift 2
for v in "$@"; do
  if [ "$v" = base ]; then L=""; else L="$PWD/little-physics-engine_b200/variants/liblpe_bh_$v.so"; fi
  echo -n "$v: "; LPE_BH_LIB="$L" python scripts/prof_step.py "$W" "$S" | sed -e "s/.*'ms_keygen'/'ms_keygen'/" -e "s/, 'pad2_.*//"
done
