#!/bin/bash
# round 2, GPU run T (1 GPU, final state): full suite, the default bench line, launch list of two bench steps
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -30) > gpurun_out/r2t_pytest.log
python bench.py > gpurun_out/r2t_bench.json 2> gpurun_out/r2t_bench.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2t_launches.csv python bench.py --steps 2 --warmup 1 --no-variants --no-cpu-baseline > gpurun_out/r2t_bench_under_ncu.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2t_smoke.log 2>&1
tail -6 gpurun_out/r2t_pytest.log | cut -c1-200; cut -c1-400 gpurun_out/r2t_bench.json; tail -n 3 gpurun_out/r2t_bench.err; tail -n 2 gpurun_out/r2t_smoke.log
