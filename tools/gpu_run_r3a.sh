set -x
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "one_to_n or fixtures or synthetic_small or full_size_c3 or ball_dilation_medium or big_regions or half_scale or series or custom_slice or axis_restriction" > gpurun_out/r3a_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r3a_tests.log
timeout 120 python -m pytest tests/test_gpu_determinism.py -x -q -m gpu > gpurun_out/r3a_tests_det.log 2>&1; echo "rc=$?" >> gpurun_out/r3a_tests_det.log
timeout 100 python bench.py --no-extra > gpurun_out/r3a_bench.json 2> gpurun_out/r3a_bench.err
MCI_SPLIT_GENERIC=1 timeout 100 python bench.py --no-extra > gpurun_out/r3a_bench_generic.json 2> gpurun_out/r3a_bench_generic.err
tail -3 gpurun_out/r3a_tests.log gpurun_out/r3a_tests_det.log
python - <<'P'
import json
for f in ("r3a_bench","r3a_bench_generic"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["e2e"]["ms_per_step"], {k:round(v["ms_per_step"],4) for k,v in d["kernels"].items()})
    except Exception as e: print(f, "ERR", e)
P
