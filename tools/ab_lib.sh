#!/bin/bash
# Same-box A/B of two builds of the library (ab/lib_old.so, ab/lib_new.so; both git-ignored):  bash tools/ab_lib.sh
# Alternates the variants twice over a few ncu_targets.py microbenchmarks and prints mean launch times.
T=${@:-gemm_fwd_bits gemm_fwd gemm_res gemm_dgrad_bits}
for i in 1 2; do
  for v in old new; do
    cp ab/lib_$v.so safevla_b200/libsafevla_b200.so
    for t in $T; do
      echo -n "$v $t: "
      python tools/ncu_targets.py $t 24 | tail -1 | python -c "
import sys,ast; l=sys.stdin.read(); xs=ast.literal_eval(l[l.index('['):]); xs=xs[4:]; print(round(sum(xs)/len(xs)*1000,1),'us')"
    done
  done
done
cp ab/lib_new.so safevla_b200/libsafevla_b200.so
