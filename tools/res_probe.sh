#!/bin/bash
# resident flow with / without the cluster-scope release + fence per layer (A/B of two libraries on one box)
timeout 300 python -m pytest tests/test_sampler_gpu.py -x -q -k "resident_mode or single_and_pair or interleaved or full_chain" 2>&1 | tail -2
for lib in default libsdrm_resold.so default libsdrm_resold.so; do
  if [ "$lib" = default ]; then unset SDRM_B200_LIB; else export SDRM_B200_LIB=sdrm_b200/csrc/$lib; fi
  echo "== $lib"
  python tools/shape_probe.py 340 490 3125 78 1 100000 2>&1 | grep SHAPE | head -1
  python tools/shape_probe.py 400 550 729 43 0 60000 2>&1 | grep SHAPE | head -1
  python tools/shape_probe.py 200 256 3000 60 3 12000 2>&1 | grep SHAPE | head -1
done
