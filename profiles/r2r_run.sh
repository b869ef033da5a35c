#!/bin/bash
# round 2, GPU run R (1 GPU): bounded sub-voxel walk per DESC execution (resumable) in the pooled trace kernel
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for v in base cap2 cap3 cap4 cap3st; do
  GSB_LIB_PATH=profiles/_variants/lib_$v.so GSB_CPF_LIST=2 timeout 300 python profiles/prof_shadow.py 103 8 1024 2>&1 | grep -E "^lib|^shadow|rays/launch|per ray|pool blocks|Error|error"
done > gpurun_out/r2r_sweep.log
(GSB_LIB_PATH=profiles/_variants/lib_cap3.so timeout 600 python -m pytest tests/test_shade_gpu.py tests/test_pipeline_gpu.py -m gpu -q --tb=short 2>&1 | tail -5) > gpurun_out/r2r_cap_parity.log
grep -E "^lib|trace_ms|per ray|pool blocks" gpurun_out/r2r_sweep.log; tail -3 gpurun_out/r2r_cap_parity.log
