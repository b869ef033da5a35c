#!/bin/bash
# round 2, GPU run U (1 GPU): the ray list taken in per-warp blocks of consecutive rays (L1 locality of the first steps of a ray)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for v in blk32 blk128 blk256 blk512 blk2048 blk8192; do
  GSB_LIB_PATH=profiles/_variants/lib_$v.so GSB_CPF_LIST=2 timeout 300 python profiles/prof_shadow.py 103 8 1024 2>&1 | grep -E "^lib|^shadow|Error|error"
done > gpurun_out/r2u_sweep.log
(timeout 600 python -m pytest tests/test_shade_gpu.py tests/test_pipeline_gpu.py -m gpu -q --tb=short 2>&1 | tail -5) > gpurun_out/r2u_parity.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-variants > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err
grep -E "^lib|trace_ms" gpurun_out/r2u_sweep.log; tail -3 gpurun_out/r2u_parity.log; cut -c1-260 gpurun_out/r2u_bench.json
