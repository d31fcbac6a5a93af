# end-of-round validation: full GPU suite, the default bench line, the reference arm, the ncu launch list
set -x
timeout 420 python -m pytest tests -x -q -m gpu > gpurun_out/r3f_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r3f_tests.log
timeout 150 python bench.py > gpurun_out/r3f_bench.json 2> gpurun_out/r3f_bench.err; echo "bench rc=$?"
timeout 60 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r3f_ref.json 2> gpurun_out/r3f_ref.err
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02c_launches.csv python bench.py --steps 2 --warmup 3 --no-extra > gpurun_out/r3f_ncu.log 2>&1; echo "ncu rc=$?"
tail -n 3 gpurun_out/r3f_tests.log
python - <<'P'
import json
try:
    d=json.loads(open("gpurun_out/r3f_bench.json").read().strip().splitlines()[-1])
    print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["gpu_launches_per_step"], d["roofline"]["frac"])
    for w in d["extra"]["workloads"]: print(w["workload"][:3], w["ms_per_step"], w["e2e"]["ms_per_step"], w["bit_exact"], w["gpu_launches_per_step"])
except Exception as e: print("ERR", e)
P
