#!/bin/bash
# compute-sanitizer over small GPU tests that touch the reworked emission (cp.async staging into the idle operand rings),
# the sized prunes, the merge pre-filter and the dependent launches.  Logs land in gpurun_out/.
mkdir -p gpurun_out
SEL="test_query_group_sizes and cosine and 5 or test_sampled_threshold_with_row_mask or test_large_batches_on_cta_pairs and 300-10-cosine or test_exchange_two_ranks_on_one_device and cosine or test_config1_correctness_reference"
timeout -k 5 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -k "$SEL" > gpurun_out/sanitizer_memcheck_r2b.log 2>&1; echo memcheck rc=$?
timeout -k 5 400 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -q -k "test_query_group_sizes and cosine and 5 or test_config1_correctness_reference" > gpurun_out/sanitizer_racecheck_r2b.log 2>&1; echo racecheck rc=$?
tail -n 4 gpurun_out/sanitizer_memcheck_r2b.log; tail -n 4 gpurun_out/sanitizer_racecheck_r2b.log
