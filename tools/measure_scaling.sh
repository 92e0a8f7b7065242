#!/bin/bash
# 1 / 2 / 4 / 8-GPU bench lines on ONE 8-GPU box (gpurun --gpus 8 -- 'bash tools/measure_scaling.sh r02').
# Launched the way the driver launches bench.py; --no-cpu: the CPU baseline leg is the N=1 default run's business.
TAG=${1:-r02}
O=gpurun_out
mkdir -p $O
python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu > $O/${TAG}_bench_c3_1gpu_8gpu_box.json 2> $O/${TAG}_1gpu.err
for n in 2 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 20 --warmup 3 --no-cpu > $O/${TAG}_bench_c3_${n}gpu.json 2> $O/${TAG}_${n}gpu.err
done
python - <<PY
import json
for n in (1, 2, 4, 8):
    f = "$O/${TAG}_bench_c3_%s.json" % ("1gpu_8gpu_box" if n == 1 else "%dgpu" % n)
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        c5 = d.get("c5") or {}
        print(n, "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "host", d.get("host_placement"))
        print("   c5", c5.get("value"), c5.get("step_ms"), c5.get("step_ms_median"), c5.get("scatter_ms"), c5.get("gather_ms"), c5.get("full_feature_scatter_ms"))
    except Exception as e:
        print(n, "ERR", e)
PY
