#!/bin/bash
# round 2, GPU run H (1 GPU): full suite, per-kernel ncu of the extraction, launch list of two bench steps, the full default bench line,
# the reference arm
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40) > gpurun_out/r2h_pytest.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_(tet|edge|occ|vertex|bwd|scan|case|raw|quads|dual|cut|boundary)' -f -o gpurun_out/r2h_extract python profiles/prof_extract.py > gpurun_out/r2h_ncu_extract.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2h_launches.csv python bench.py --steps 2 --warmup 1 --no-variants --no-cpu-baseline > gpurun_out/r2h_bench_under_ncu.log 2>&1
python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err
(time python bench.py --impl reference --steps 20 --warmup 5) > gpurun_out/r2h_bench_reference.json 2> gpurun_out/r2h_bench_reference.err
tail -8 gpurun_out/r2h_pytest.log | cut -c1-200; cut -c1-400 gpurun_out/r2h_bench.json; tail -3 gpurun_out/r2h_bench.err; cut -c1-600 gpurun_out/r2h_bench_reference.json; tail -4 gpurun_out/r2h_bench_reference.err
