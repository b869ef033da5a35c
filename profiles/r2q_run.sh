#!/bin/bash
# round 2, GPU run Q (8 GPUs): bench lines at 8 and 4 ranks with the forward shading dealt out over the ranks
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 8 --steps 5 --warmup 3 --no-variants --no-cpu-baseline > gpurun_out/r2q_bench_8gpu.json 2> gpurun_out/r2q_bench8.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29632 bench.py --gpus 4 --steps 5 --warmup 3 --no-variants --no-cpu-baseline > gpurun_out/r2q_bench_4gpu.json 2> gpurun_out/r2q_bench4.err
python - <<'PY'
import json
for n in (8, 4):
    try:
        txt=open(f'gpurun_out/r2q_bench_{n}gpu.json').read()
        d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
        print(d['n_gpus'], d['ms_per_step'], json.dumps(d['stages_ms_max_over_ranks']))
    except Exception as e:
        print(n, "failed", e)
PY
tail -n 3 gpurun_out/r2q_bench8.err gpurun_out/r2q_bench4.err
