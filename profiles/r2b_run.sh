#!/bin/bash
# round 2, GPU run B: trace kernel v2 (lean step, in-cell fine walk, grid by value): parity tests, probe, counters, knob sweep, bench
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_shade_gpu.py tests/test_pipeline_gpu.py tests/test_flex_gpu.py -m gpu -q --tb=short 2>&1 | tail -70) > gpurun_out/r2b_pytest.log
GSB_CPF_LIST=1,2,4 python profiles/prof_shadow.py 103 8 1024 > gpurun_out/r2b_probe_base.log 2>&1
GSB_LIB_PATH=profiles/_variants/lib_stats.so GSB_CPF_LIST=2,4 python profiles/prof_shadow.py 103 8 1024 > gpurun_out/r2b_probe_stats.log 2>&1
for v in s2 s6 s8 vt4 vt12 vt16 vt20 ms6 ms18 rf22 rf30 b5 b6 bt2 bt4 t128 s8vt16; do
  GSB_LIB_PATH=profiles/_variants/lib_$v.so GSB_CPF_LIST=2 python profiles/prof_shadow.py 103 8 1024 2>&1 | grep -E "^lib|^shadow|rays/launch|Error|error"
done > gpurun_out/r2b_sweep.log
python profiles/prof_shadow.py 103 8 1024 sphere > gpurun_out/r2b_probe_sphere.log 2>&1
python bench.py --steps 5 --warmup 3 --no-variants --no-cpu-baseline > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
tail -3 gpurun_out/r2b_pytest.log; tail -8 gpurun_out/r2b_probe_base.log; cat gpurun_out/r2b_sweep.log
