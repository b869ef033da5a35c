#!/bin/bash
# round 2, GPU run S (1 GPU): rays gathered per lane in GEN (ray order seen by the trace kernel), register budgets of FWD / BWD
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for v in base gen2 gen8 gen16; do
  GSB_LIB_PATH=profiles/_variants/lib_$v.so GSB_CPF_LIST=2 timeout 300 python profiles/prof_shadow.py 103 8 1024 2>&1 | grep -E "^lib|^shadow|no-shadow|Error|error"
done > gpurun_out/r2s_sweep.log
for v in base bwd5 fwd8; do
  echo "lib $v"; GSB_LIB_PATH=profiles/_variants/lib_$v.so python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variants 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['roofline']['other_kernels']['env_shade']
print(d['ms_per_step'], 'bwd alone', e['ms'], 'fwd alone', e['fwd']['ms'], 'backward_total', d['stages_ms_max_over_ranks']['backward_total'], 'env_shade_fwd', d['stages_ms_max_over_ranks']['env_shade_fwd(gen+trace+shade)'])"
done > gpurun_out/r2s_bench.log 2>&1
grep -E "^lib|trace_ms|no-shadow" gpurun_out/r2s_sweep.log; cat gpurun_out/r2s_bench.log
