#!/bin/bash
# round 2, GPU run I (1 GPU): pooled-context trace kernel (k_trace_pool) against the 2-context kernel, its parity, the full suite, a short bench
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for v in c2b8 c2v12 p2 p2s2 p2s6 p2s8 p2bt4 p2r8 p2r24 p2bias4 p3b6 p3b5 p4b4 p2s6bt4; do
  GSB_LIB_PATH=profiles/_variants/lib_$v.so GSB_CPF_LIST=2 timeout 300 python profiles/prof_shadow.py 103 8 1024 2>&1 | grep -E "^lib|^shadow|rays/launch|per ray|Error|error"
done > gpurun_out/r2i_sweep.log
(GSB_LIB_PATH=profiles/_variants/lib_p2.so timeout 600 python -m pytest tests/test_shade_gpu.py tests/test_pipeline_gpu.py -m gpu -q --tb=short 2>&1 | tail -15) > gpurun_out/r2i_pool_parity.log
(timeout 1200 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40) > gpurun_out/r2i_pytest.log
timeout 300 python profiles/prof_extract.py > gpurun_out/r2i_extract.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
grep -E "^lib|trace_ms" gpurun_out/r2i_sweep.log; tail -6 gpurun_out/r2i_pool_parity.log; tail -12 gpurun_out/r2i_pytest.log | cut -c1-220; tail -12 gpurun_out/r2i_extract.log; cut -c1-300 gpurun_out/r2i_bench.json; tail -3 gpurun_out/r2i_bench.err
