#!/bin/bash
# development A/B: run a script against several builds of the product library (gpurun_in_<tag>.so at the repo root)
cp quantizedattention_b200/libqattn.so /tmp/keep.so
for f in gpurun_in_*.so; do
  cp "$f" quantizedattention_b200/libqattn.so
  echo "== $f"; timeout 300 python "$@" 2>&1 | grep -E -A4 "ms_call|TOPS|ms_kernel|passed|failed|Error|error|checksums" | head -40
done
cp /tmp/keep.so quantizedattention_b200/libqattn.so
