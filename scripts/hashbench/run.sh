#!/bin/bash
cd "$(dirname "$0")"
grep -m1 "model name" /proc/cpuinfo; nproc
for fl in "-O2" "-O3" "-O3 -mbmi2" "-O3 -march=haswell" "-O3 -march=native"; do
  g++ $fl -std=c++17 -o /tmp/hb_t hb.cpp 2>/dev/null && echo "== $fl" && /tmp/hb_t | tail -2
done
