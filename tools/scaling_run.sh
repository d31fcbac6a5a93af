#!/bin/bash
# bench.py at 1, 2, 4, 8 GPUs of one box, the way the driver launches it (weak scaling: one C3 volume per GPU).
# Usage (on the GPU box, 8 GPUs visible): bash tools/scaling_run.sh [steps]
STEPS=${1:-20}
mkdir -p gpurun_out
python bench.py --gpus 1 --steps $STEPS --warmup 3 > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
for N in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) \
    bench.py --gpus $N --steps $STEPS --warmup 3 --no-check > gpurun_out/scale_n$N.json 2> gpurun_out/scale_n$N.err
done
python - <<'PY'
import json
base = None
for n in (1, 2, 4, 8):
    try:
        d = json.loads([l for l in open("gpurun_out/scale_n%d.json" % n).read().splitlines() if l.startswith("{")][-1])
    except Exception as e:
        print(n, "failed", e)
        continue
    base = base or d["value"]
    print("N=%d value %.0f Mvoxel/s (%.3f ms/step, x%.2f)  e2e %.0f Mvoxel/s (%.2f ms)  one-at-a-time %.0f  clocks %s" % (
        n, d["value"], d["ms_per_step"], d["value"] / base, d["e2e"]["value"], d["e2e"]["ms_per_step"],
        d["e2e"]["one_volume_at_a_time"]["value"], d["clocks"]))
PY
