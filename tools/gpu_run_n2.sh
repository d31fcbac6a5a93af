set -x
timeout 85 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02c_bench_n2.json 2> gpurun_out/r02c_bench_n2.err; echo "rc=$?"
tail -c 600 gpurun_out/r02c_bench_n2.err
python - <<'P'
import json
try:
    d=json.loads([l for l in open("gpurun_out/r02c_bench_n2.json").read().splitlines() if l.startswith("{")][-1])
    print(d["value"], d["ms_per_step"], d["e2e"]["value"]); print(json.dumps(d.get("sharded"))[:1500])
except Exception as e: print("ERR", e)
P
