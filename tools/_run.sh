for v in b200 nc256; do
  export SDRM_B200_LIB=$PWD/sdrm_b200/csrc/libsdrm_$v.so
  echo "== variant $v"
  bash tools/quick_bench.sh 37888 "2:37888 2:256"
done
unset SDRM_B200_LIB
timeout 900 python -m pytest tests/test_sampler_gpu.py -x -q 2>&1 | tail -3
