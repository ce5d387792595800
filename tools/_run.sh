timeout 600 python -m pytest tests/test_sampler_gpu.py tests/test_probe_gpu.py -x -q 2>&1 | tail -3
bash tools/quick_bench.sh 37888 "2:37888 2:256 1:18944"
