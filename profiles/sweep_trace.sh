# sweep of the trace kernel's build-time knobs on the full-size random-SDF soup (run on the GPU box)
run() { touch gshell_b200/csrc/occluder.cu; python -m gshell_b200.build > /dev/null 2>&1; echo "$1: $(python profiles/prof_shadow.py 103 8 1024 2>&1 | grep '^shadow' | cut -c1-40)"; }
python -m pytest tests/test_shade_gpu.py -m gpu -q 2>&1 | tail -1
for b in 2 3 4; do for st in 2 4; do
  GSB_TRACE_BATCH=$b GSB_TRACE_STEPS=$st run "BATCH $b STEPS $st"
done; done
export GSB_TRACE_THREADS=128 GSB_TRACE_MIN_BLOCKS=6 GSB_TRACE_BLOCKS=6
GSB_TRACE_BATCH=4 GSB_TRACE_STEPS=4 run "BATCH 4 STEPS 4 128x6"
