#!/bin/bash
# round 2, GPU run N (8 GPUs): sharded-step exactness test on NCCL (2 ranks), bench lines at 8 and 4 ranks with the stage breakdown
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2n_gpus.txt
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 5 --warmup 3 --no-variants --no-cpu-baseline > gpurun_out/r2n_bench_8gpu.json 2> gpurun_out/r2n_bench8.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 4 --steps 5 --warmup 3 --no-variants --no-cpu-baseline > gpurun_out/r2n_bench_4gpu.json 2> gpurun_out/r2n_bench4.err
(timeout 400 python -m pytest tests/test_multigpu_gpu.py -m gpu -q --tb=short -s 2>&1 | tail -30) > gpurun_out/r2n_pytest.log
tail -6 gpurun_out/r2n_pytest.log; python - <<'PY'
import json
for n in (8, 4):
    try:
        d=json.load(open(f'gpurun_out/r2n_bench_{n}gpu.json'))
        print(d['n_gpus'], d['ms_per_step'], json.dumps(d['stages_ms_max_over_ranks']))
    except Exception as e:
        print(n, "failed", e)
PY
tail -3 gpurun_out/r2n_bench8.err gpurun_out/r2n_bench4.err
