#!/bin/bash
# round 2, GPU run G (2 GPUs): sharded-step exactness test on NCCL, 2-rank bench line with the stage breakdown
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2g_gpus.txt
(timeout 600 python -m pytest tests/test_multigpu_gpu.py -m gpu -q --tb=short -s 2>&1 | tail -30) > gpurun_out/r2g_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29601 bench.py --gpus 2 --steps 5 --warmup 3 --no-variants > gpurun_out/r2g_bench_2gpu.json 2> gpurun_out/r2g_bench.err
tail -12 gpurun_out/r2g_pytest.log; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2g_bench_2gpu.json'))
print(d['n_gpus'], d['ms_per_step'], json.dumps(d['stages_ms_max_over_ranks']))
PY
tail -3 gpurun_out/r2g_bench.err
