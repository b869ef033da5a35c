#!/bin/bash
# round 2, GPU run M (1 GPU): cooperative TEST block (several lanes per slot) of the pooled trace kernel
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for v in base coop4 coop8c2 coop8c4 coop16 coop8st coop8s4 coop8rf12; do
  GSB_LIB_PATH=profiles/_variants/lib_$v.so GSB_CPF_LIST=2 timeout 300 python profiles/prof_shadow.py 103 8 1024 2>&1 | grep -E "^lib|^shadow|rays/launch|per ray|pool blocks|Error|error"
done > gpurun_out/r2m_sweep.log
(GSB_LIB_PATH=profiles/_variants/lib_coop8c2.so timeout 600 python -m pytest tests/test_shade_gpu.py tests/test_pipeline_gpu.py -m gpu -q --tb=short 2>&1 | tail -8) > gpurun_out/r2m_coop_parity.log
grep -E "^lib|trace_ms|per ray|pool blocks" gpurun_out/r2m_sweep.log; tail -4 gpurun_out/r2m_coop_parity.log
