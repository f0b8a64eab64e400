#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_transformer_gpu.py -m gpu -q > gpurun_out/pytest_batch.txt 2>&1; echo "bst pytest rc=$?"
tail -2 gpurun_out/pytest_batch.txt; grep -E "^(FAILED|ERROR)" gpurun_out/pytest_batch.txt | head -10
timeout 300 python tools/bench_bst.py 2>&1 | cut -c1-170 | tee gpurun_out/bench_bst_r2.txt
timeout 600 python -m pytest tests/test_tc_gpu.py -m gpu -q -x -k "pair_tiles" > gpurun_out/pytest_pair.txt 2>&1; echo "pair pytest rc=$?"
tail -2 gpurun_out/pytest_pair.txt; grep -E "^(FAILED|ERROR)|Error|error:" gpurun_out/pytest_pair.txt | head -10
for e in "BSMM_PAIR_TILES=0" "BSMM_PAIR_TILES=1"; do echo "== $e"; env $e timeout 200 python tools/time_ops.py 0.05 0.10 0.25 2>&1 | sed 's/| updat.*//'; done | tee gpurun_out/time_pair.txt
