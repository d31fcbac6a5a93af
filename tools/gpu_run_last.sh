set -x
timeout 40 python -m pytest tests/test_cpp_filter.py tests/test_upstream_baselines.py tests/test_gpu_sharded.py -x -q -m gpu -k "not three_shards and not combine_by" > gpurun_out/r3g_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r3g_tests.log
tail -n 4 gpurun_out/r3g_tests.log
