#!/bin/bash
# A/B timing of two builds of the library on the same GPU box: tools/ab.sh [members]
for rep in 1 2; do for v in a b; do echo -n "$v: "; SPEEDY_B200_LIB=$PWD/speedy.f90_b200/libspeedy_b200_$v.so timeout 200 python tools/ktime.py ${1:-1} 2>&1 | grep members | cut -c1-75; done; done
