#!/bin/bash
# round 2, GPU run F: full -m gpu suite, bench with the per-stage breakdown, ncu of the env_shade kernels and of the default trace kernel
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -60) > gpurun_out/r2f_pytest.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_env_shade -c 3 -f -o gpurun_out/r2f_env_shade python profiles/prof_env_shade.py > gpurun_out/r2f_ncu_env.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 2 -c 1 -f -o gpurun_out/r2f_trace python profiles/prof_shadow.py 103 8 1024 > gpurun_out/r2f_ncu_trace.log 2>&1
tail -40 gpurun_out/r2f_pytest.log | grep -E "FAILED|passed|failed|^E  " | cut -c1-200; python - <<'PY'
import json
d=json.load(open('gpurun_out/r2f_bench.json'))
print(d['ms_per_step'], d['gpu_launches'], json.dumps(d['stages_ms_max_over_ranks']), json.dumps(d['variants'])[:600])
r=d['roofline']; print({k:r.get(k) for k in ('ms_total','launches','share_of_step','rays_per_s','frac')}); print(json.dumps(r['other_kernels'])[:900])
PY
tail -3 gpurun_out/r2f_bench.err
