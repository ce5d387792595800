set -x
CMD="python bench.py --rows 18944 --steps 2 --warmup 1 --no-cpu --no-e2e"
$CMD > gpurun_out/plain_r01.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv $CMD > gpurun_out/ncu_list.log 2>&1
tail -2 gpurun_out/ncu_list.log
$CMD > gpurun_out/plain_r01.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sdrm_layer_engine -s 1 -c 1 -f -o gpurun_out/prof_r01 $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
CMD2="python tools/bench_sparsify.py 125000 20000"
$CMD2 > gpurun_out/plain_k4.log 2>&1 && \
ncu --set full --clock-control none -k regex:key_hist\|threshold_pack -c 6 -f -o gpurun_out/prof_k4_r01 $CMD2 > gpurun_out/ncu_k4.log 2>&1
tail -2 gpurun_out/ncu_k4.log
