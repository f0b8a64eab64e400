#!/bin/bash
# compute-sanitizer evidence for the hand-rolled mbarrier / TMA / tcgen05 protocols (small parity cases; kernels run 10-100x slower,
# so the in-kernel wait bound is raised).
mkdir -p gpurun_out
export BSMM_WAIT_TIMEOUT_MS=120000
T=tests/test_tc_gpu.py
IDS="$T::test_tc_xprop_matches_oracle[case0-dtype0-1] $T::test_tc_xprop_matches_oracle[case0-dtype0-0] $T::test_tc_xprop_matches_oracle[case9-dtype0-1] $T::test_tc_updat_matches_oracle[case0-dtype0-1] $T::test_tc_updat_matches_oracle[case6-dtype0-1] $T::test_tc_xprop_pair_tiles_match_oracle[case0-dtype0-1] $T::test_bs8_runs_padded_on_tcgen05[dtype0-1] $T::test_tc_xprop2_variants_match_oracle[case0-dtype0-1-2] $T::test_tc_xprop2_variants_match_oracle[case0-dtype0-1-3]"
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest $IDS -m gpu -q -x > gpurun_out/sanitizer_memcheck_tc.txt 2>&1; echo "memcheck tc rc=$?"
tail -6 gpurun_out/sanitizer_memcheck_tc.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 99 --print-limit 20 python -m pytest "$T::test_tc_xprop_matches_oracle[case0-dtype0-1]" "$T::test_tc_updat_matches_oracle[case0-dtype0-1]" -m gpu -q -x > gpurun_out/sanitizer_racecheck_tc.txt 2>&1; echo "racecheck rc=$?"
tail -8 gpurun_out/sanitizer_racecheck_tc.txt
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 99 --print-limit 20 python -m pytest "$T::test_tc_xprop_matches_oracle[case0-dtype0-1]" "$T::test_tc_updat_matches_oracle[case0-dtype0-1]" -m gpu -q -x > gpurun_out/sanitizer_synccheck_tc.txt 2>&1; echo "synccheck rc=$?"
tail -5 gpurun_out/sanitizer_synccheck_tc.txt
