#!/bin/bash
# round 2, GPU run C: block order S->D->T, several ray contexts per lane (shared-memory state): parity + sweep
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for v in base c2 c3 c4 c6 c3s2 c3s8 c3v c3v1 c3r2 c3r16 c3t256 c3b4 c3stats; do
  GSB_LIB_PATH=profiles/_variants/lib_$v.so GSB_CPF_LIST=2 python profiles/prof_shadow.py 103 8 1024 2>&1 | grep -E "^lib|^shadow|rays/launch|per ray|Error|error"
done > gpurun_out/r2c_sweep.log
(GSB_LIB_PATH=profiles/_variants/lib_c3.so timeout 900 python -m pytest tests/test_shade_gpu.py tests/test_pipeline_gpu.py tests/test_flex_gpu.py -m gpu -q --tb=short 2>&1 | tail -30) > gpurun_out/r2c_pytest_c3.log
GSB_LIB_PATH=profiles/_variants/lib_c3.so python profiles/prof_shadow.py 103 8 1024 sphere > gpurun_out/r2c_probe_sphere_c3.log 2>&1
timeout 600 env GSB_LIB_PATH=profiles/_variants/lib_c3.so ncu --set full --clock-control none --import-source on -k regex:k_trace -s 2 -c 1 -f -o gpurun_out/r2c_trace_c3 python profiles/prof_shadow.py 103 8 1024 > gpurun_out/r2c_ncu.log 2>&1
cat gpurun_out/r2c_sweep.log; tail -5 gpurun_out/r2c_pytest_c3.log
