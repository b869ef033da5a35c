#!/bin/bash
# round 2, GPU run O (2 GPUs): sharded-step tests (plumbing; dealt forward shading), 2-rank bench line
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_multigpu_gpu.py tests/test_shade_gpu.py -m gpu -q --tb=short -s 2>&1 | tail -40) > gpurun_out/r2o_pytest.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --steps 5 --warmup 3 --no-variants --no-cpu-baseline > gpurun_out/r2o_bench_2gpu.json 2> gpurun_out/r2o_bench.err
grep -E "max err|passed|failed|Error" gpurun_out/r2o_pytest.log | tail -24; python - <<'PY'
import json
txt=open('gpurun_out/r2o_bench_2gpu.json').read()
d=json.loads([l for l in txt.splitlines() if l.startswith('{')][-1])
print(d['n_gpus'], d['ms_per_step'], json.dumps(d['stages_ms_max_over_ranks']))
PY
tail -n 3 gpurun_out/r2o_bench.err
