#!/bin/bash
# same-box A/B of two library builds on the attention sweep (ab/lib_old.so vs ab/lib_new.so)
for i in 1 2; do for v in old new; do cp ab/lib_$v.so safevla_b200/libsafevla_b200.so; echo "== $v"; python tools/attn_sweep.py > /tmp/a.txt 2>&1; grep "mode" /tmp/a.txt | head -${1:-3}; done; done
cp ab/lib_new.so safevla_b200/libsafevla_b200.so
