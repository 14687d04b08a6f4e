#!/bin/bash
# bench.py at N GPUs as the driver launches it: tools/run_r2_scale.sh N
N=$1; mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_n$N.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("n_gpus","value","ms_per_step")}, "e2e", d["e2e"]["value"], "dense e2e", d["e2e"]["dense_uint8_frames"]["value"], d["e2e"]["h2d_bytes_per_step"])
PY
tail -2 gpurun_out/r2_bench_n$N.err
