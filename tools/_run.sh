timeout 900 python -m pytest tests/test_training_gpu.py tests/test_abi.py -x -q 2>&1 | tail -5
rm -f gpurun_out/e2e_ours_ml100k.jsonl
timeout 1500 python tools/our_e2e.py 5 1 gpurun_out/e2e_ours_ml100k.jsonl 2>&1 | grep impl | cut -c1-330
