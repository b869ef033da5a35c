#!/bin/bash
# round 2, GPU run D: occupancy / prefetch sweep of the multi-context trace kernel; parity of the fused tick() loss kernels
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for v in c2b8 c2b9 c2b10 c1b12 c1b10 c2t64 c2pb c2pc c2pr1 c2pr2 c2pall c2b9bt2 c2b9bt4 c2b9s2 c2b9s6 c3b7; do
  GSB_LIB_PATH=profiles/_variants/lib_$v.so GSB_CPF_LIST=2 python profiles/prof_shadow.py 103 8 1024 2>&1 | grep -E "^lib|^shadow|rays/launch|per ray|Error|error"
done > gpurun_out/r2d_sweep.log
(timeout 900 python -m pytest tests/test_glue_gpu.py tests/test_pipeline_gpu.py tests/test_flex_gpu.py -m gpu -q --tb=short 2>&1 | tail -40) > gpurun_out/r2d_pytest.log
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
cat gpurun_out/r2d_sweep.log; tail -25 gpurun_out/r2d_pytest.log; cut -c1-300 gpurun_out/r2d_bench.json; tail -3 gpurun_out/r2d_bench.err
