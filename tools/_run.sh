SDRM_HIST_MERGE=0 python tools/bench_sparsify.py 125000 20000 2>&1 | grep pass1
