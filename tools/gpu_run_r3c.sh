set -x
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fixtures or synthetic_small or full_size_c3 or ball_dilation_medium or slice_detection or label_restriction or axis_restriction or many_shapes or int16_and_uint8 or negative or custom_slice or series or one_to_n or empty" > gpurun_out/r3c_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r3c_tests.log
timeout 100 python bench.py --no-extra > gpurun_out/r3c_bench.json 2> gpurun_out/r3c_bench.err
tail -n 3 gpurun_out/r3c_tests.log
python - <<'P'
import json
for f in ("r3c_bench",):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches_per_step"], {k:round(v["ms_per_step"],4) for k,v in d["kernels"].items()})
    except Exception as e: print(f, "ERR", e)
P
