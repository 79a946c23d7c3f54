#!/bin/bash
# compute-sanitizer over the smoke path and two small parity tests (memcheck), shared-memory racecheck over smoke
out=gpurun_out; mkdir -p $out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > $out/san_memcheck_smoke.log 2>&1; echo "memcheck smoke exit $?"
tail -4 $out/san_memcheck_smoke.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -m pytest tests -m gpu -x -q -k "test_update_tsdf_interpolated_chains_replay_rounds or test_preprocess_scan_device_matches_oracle or test_track_scan or test_g2_shift or test_fused_peer_registration_one_process" > $out/san_memcheck_tests.log 2>&1; echo "memcheck tests exit $?"
tail -5 $out/san_memcheck_tests.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" > $out/san_racecheck_smoke.log 2>&1; echo "racecheck smoke exit $?"
tail -4 $out/san_racecheck_smoke.log
