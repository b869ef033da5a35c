#!/bin/bash
# round 2, GPU run L (1 GPU): L1 no-allocate knobs for the record streams of the pooled trace kernel; ncu of the default
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for v in base natri narec naboth s5 s5rf12 s5natri; do
  GSB_LIB_PATH=profiles/_variants/lib_$v.so GSB_CPF_LIST=2 timeout 300 python profiles/prof_shadow.py 103 8 1024 2>&1 | grep -E "^lib|^shadow|rays/launch|per ray|pool blocks|Error|error"
done > gpurun_out/r2l_sweep.log
GSB_LIB_PATH=profiles/_variants/lib_base.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 2 -c 1 -f -o gpurun_out/r2l_trace python profiles/prof_shadow.py 103 8 1024 > gpurun_out/r2l_ncu_trace.log 2>&1
grep -E "^lib|trace_ms" gpurun_out/r2l_sweep.log; tail -3 gpurun_out/r2l_ncu_trace.log
