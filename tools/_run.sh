set -x
for c in 2 8; do
  r=18944; [ $c = 8 ] && r=15360
  CMD="python bench.py --rows $r --steps 1 --warmup 1 --no-cpu --no-e2e --cluster $c"
  $CMD > gpurun_out/plain_c$c.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:sdrm_layer_engine -s 1 -c 1 -f -o gpurun_out/prof_c$c $CMD > gpurun_out/ncu_c$c.log 2>&1
  tail -2 gpurun_out/ncu_c$c.log
done
