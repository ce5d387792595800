#!/bin/bash
# A/B of library variants on one box: tools/ab_libs2.sh workload lib1 lib2 ...   (default library: "default")
WL=$1; shift
for rep in 1 2; do
for lib in "$@"; do
  if [ "$lib" = default ]; then unset SDRM_B200_LIB; else export SDRM_B200_LIB=sdrm_b200/csrc/$lib; fi
  timeout 300 python bench.py --workload $WL --steps 20 --warmup 3 --no-cpu --no-secondary --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib', round(d['ms_per_step'],3), 'ms', round(d['value']), 'users/s')"
done; done
