timeout 900 python -m pytest tests/test_sampler_gpu.py -x -q 2>&1 | tail -3
for w in cfg1 cfg2 cfg3 cfg4; do
  python bench.py --workload $w --steps 10 --warmup 3 --e2e-steps 5 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err || tail -3 gpurun_out/bench_$w.err
  python - $w <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/bench_{sys.argv[1]}.json"))
print("WL", sys.argv[1], "users/s", round(d["value"]), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 4), "cpu", round(d["cpu_baseline"]["value"], 1), "cluster", d["cluster"])
PY
done
