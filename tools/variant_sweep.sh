#!/bin/bash
# run tools/quick_perf.py against every engine variant in build/variants (experiments only)
for lib in build/variants/*.so; do
  echo "== $lib"
  LF_ENGINE_LIB=$PWD/$lib python tools/quick_perf.py ${1:-2e6} ${2:-1024} ${3:-free} ${4:-f64} 2>&1 | grep -A1 "call 4" | cut -c1-150
done
