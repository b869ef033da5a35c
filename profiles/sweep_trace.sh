# sweep of the trace kernel's build-time knobs on the full-size random-SDF soup (run on the GPU box)
run() { touch gshell_b200/csrc/occluder.cu; python -m gshell_b200.build > /dev/null 2>&1; echo "$1: $(python profiles/prof_shadow.py 103 8 1024 2>&1 | grep '^shadow' | cut -c1-40)"; }
run "base"
export GSB_TRACE_PF=1
run "PF"
python -m pytest tests/test_shade_gpu.py -m gpu -q 2>&1 | tail -1
GSB_TRACE_EARLY=1 run "PF EARLY"
GSB_TRACE_STEPS=1 run "PF STEPS 1"
GSB_TRACE_STEPS=3 run "PF STEPS 3"
