#!/bin/bash
# round 2, GPU run J (1 GPU): pooled trace kernel -- block statistics, batch / rounds / cells-per-face sweep, ncu of the best candidate
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for v in p2bt3st p2bt4st p2bt4 p2bt5 p2bt6 p2bt4r2 p2bt3r2 p2bt8b6 p2bt4rf28 p3bt4b6 p2bt4b7 p2bt4bias; do
  cpf=2; [ $v = p2bt4 ] && cpf=1,2,4,8
  GSB_LIB_PATH=profiles/_variants/lib_$v.so GSB_CPF_LIST=$cpf timeout 300 python profiles/prof_shadow.py 103 8 1024 2>&1 | grep -E "^lib|^shadow|rays/launch|per ray|pool blocks|Error|error"
done > gpurun_out/r2j_sweep.log
GSB_LIB_PATH=profiles/_variants/lib_p2bt4.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 2 -c 1 -f -o gpurun_out/r2j_trace python profiles/prof_shadow.py 103 8 1024 > gpurun_out/r2j_ncu_trace.log 2>&1
grep -E "^lib|trace_ms|per ray|pool blocks" gpurun_out/r2j_sweep.log; tail -3 gpurun_out/r2j_ncu_trace.log
