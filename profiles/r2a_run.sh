#!/bin/bash
# round 2, GPU run A: tests, trace probe (resolutions, counters, build-knob sweep), one ncu capture, a short bench
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/r2a_gpu.txt
(timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -40) > gpurun_out/r2a_pytest.log
GSB_CPF_LIST=2,4,8 python profiles/prof_shadow.py 103 8 1024 > gpurun_out/r2a_probe_base.log 2>&1
GSB_LIB_PATH=profiles/_variants/lib_stats.so GSB_CPF_LIST=2,4,8 python profiles/prof_shadow.py 103 8 1024 > gpurun_out/r2a_probe_stats.log 2>&1
for v in s8 s2 vt4 vt16 vd2 ms20 rf30 rf20 b5 b3 bt2 bt4; do
  GSB_LIB_PATH=profiles/_variants/lib_$v.so GSB_CPF_LIST=2 python profiles/prof_shadow.py 103 8 1024 2>&1 | grep -E "^lib|^shadow|rays/launch"
done > gpurun_out/r2a_sweep.log
python profiles/prof_shadow.py 103 8 1024 sphere > gpurun_out/r2a_probe_sphere.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace_list -s 2 -c 1 -f -o gpurun_out/r2a_trace python profiles/prof_shadow.py 103 8 1024 > gpurun_out/r2a_ncu.log 2>&1
python bench.py --steps 5 --warmup 3 --no-variants --no-cpu-baseline > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -3 gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_probe_base.log | tail -8; cat gpurun_out/r2a_bench.json | cut -c1-400
