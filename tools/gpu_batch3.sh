#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_matmul_gpu.py tests/test_transformer_gpu.py tests/test_tc_gpu.py -m gpu -q -k "bs8 or golden or softmax or transformer or bst" > gpurun_out/pytest_batch.txt 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/pytest_batch.txt; grep -E "^(FAILED|ERROR)" gpurun_out/pytest_batch.txt | head -10
timeout 300 python tools/bench_bst.py 2>&1 | grep softmax | cut -c1-170 | tee gpurun_out/bench_bst_r2b.txt
