#!/bin/bash
# round 2, GPU run P (1 GPU): full suite and a short bench after the sub-voxel kernel of the occluder build, the grouped GEN append and the pixel-id ABI
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30) > gpurun_out/r2p_pytest.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
GSB_CPF_LIST=2 timeout 300 python profiles/prof_shadow.py 103 8 1024 2>&1 | grep -E "^lib|^shadow|rays/launch|no-shadow" > gpurun_out/r2p_probe.log
timeout 300 python profiles/prof_env_shade.py > gpurun_out/r2p_env_shade.log 2>&1
tail -8 gpurun_out/r2p_pytest.log | cut -c1-200; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2p_bench.json'))
print(d['ms_per_step'], json.dumps(d['stages_ms_max_over_ranks']))
PY
tail -n 3 gpurun_out/r2p_bench.err; cat gpurun_out/r2p_probe.log; tail -5 gpurun_out/r2p_env_shade.log
