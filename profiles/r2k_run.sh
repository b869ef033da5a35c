#!/bin/bash
# round 2, GPU run K (1 GPU): pooled trace kernel on the packed-position traversal (lean cell step): sweep, block statistics, parity suite
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for v in base st s3 s5 s6 r1 bt5r2 bt3r3 bt6r1 rf12 rf24 b9 b7; do
  GSB_LIB_PATH=profiles/_variants/lib_$v.so GSB_CPF_LIST=2 timeout 300 python profiles/prof_shadow.py 103 8 1024 2>&1 | grep -E "^lib|^shadow|rays/launch|per ray|pool blocks|Error|error"
done > gpurun_out/r2k_sweep.log
(timeout 1200 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -40) > gpurun_out/r2k_pytest.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err
grep -E "^lib|trace_ms|per ray|pool blocks" gpurun_out/r2k_sweep.log; tail -8 gpurun_out/r2k_pytest.log | cut -c1-220; cut -c1-300 gpurun_out/r2k_bench.json; tail -3 gpurun_out/r2k_bench.err
