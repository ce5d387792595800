bash tools/quick_bench.sh 37888 "2:37888"
for f in 0.5 0.9; do
echo "== persist $f"
SDRM_L2_PERSIST=$f bash tools/quick_bench.sh 37888 "2:37888 8:30720"
grep sdrm gpurun_out/qb_2.err | head -2
done
