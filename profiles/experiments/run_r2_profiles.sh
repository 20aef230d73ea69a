#!/bin/bash
# Round-2 profile + sanitizer capture (run on a B200 through gpurun; outputs land in gpurun_out/, summaries are made by
# profiles/summarize.py r2 and profiles/sass_summary.py on the build host).
mkdir -p gpurun_out
B="python bench.py --no-also --no-cpu-baseline --no-parity"
timeout -k 5 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r2_headline.csv $B --steps 5 --warmup 3 > gpurun_out/r2l_ncu1.log 2>&1; echo ncu1 rc=$?
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:knn_scan_shadow -c 1 -o gpurun_out/prof_r2_shadow_headline $B --steps 1 --warmup 1 > gpurun_out/r2l_ncu2.log 2>&1; echo ncu2 rc=$?
timeout -k 5 300 ncu --set full --clock-control none -k regex:filter_finish -c 1 -o gpurun_out/prof_r2_finish_headline $B --steps 1 --warmup 1 > gpurun_out/r2l_ncu3.log 2>&1; echo ncu3 rc=$?
timeout -k 5 300 ncu --set full --clock-control none -k regex:filter_prep -c 1 -o gpurun_out/prof_r2_prep_headline $B --steps 1 --warmup 1 > gpurun_out/r2l_ncu4.log 2>&1; echo ncu4 rc=$?
# compute-sanitizer over a few small GPU tests that touch every new kernel (the reference's counterpart is `go test -race`)
SEL="test_query_group_sizes and cosine and 5 or test_group_search_best_of_chunks and cosine or test_large_batches_on_cta_pairs and 256-10-cosine or test_exchange_two_ranks_on_one_device and cosine or test_sampled_threshold_with_row_mask"
timeout -k 5 500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -q -k "$SEL" > gpurun_out/sanitizer_memcheck_r2.log 2>&1; echo memcheck rc=$?
timeout -k 5 300 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -q -k "test_query_group_sizes and cosine and 5" > gpurun_out/sanitizer_racecheck_r2.log 2>&1; echo racecheck rc=$?
tail -3 gpurun_out/sanitizer_memcheck_r2.log gpurun_out/sanitizer_racecheck_r2.log
