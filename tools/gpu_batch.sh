#!/bin/bash
# one GPU call, several checks: tcgen05 parity (incl. bs 16, CUDA graph), transformer parity (staged softmax), timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_transformer_gpu.py tests/test_baseline_configs_gpu.py -m gpu -q > gpurun_out/pytest_batch.txt 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/pytest_batch.txt; grep -E "^(FAILED|ERROR)" gpurun_out/pytest_batch.txt | head -20
timeout 300 python tools/bench_bst.py 2>&1 | cut -c1-170 | tee gpurun_out/bench_bst_r2.txt
BSMM_SOFTMAX_STAGED=0 timeout 300 python tools/bench_bst.py 2>&1 | grep softmax | cut -c1-170 | tee -a gpurun_out/bench_bst_r2.txt
timeout 300 python tools/bench_cfg4.py 2>&1 | cut -c1-420 | tee gpurun_out/bench_cfg4_r2.txt
