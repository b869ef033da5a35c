#!/bin/bash
# wrapper used in the build container: rebuild the library and every profiling variant named in $VARIANTS against the CURRENT
# sources, then hand the script to gpurun (stale variant .so files cost two sweeps in round 2)
set -e
cd "$(dirname "$0")/.."
python -m gshell_b200.build > /dev/null
rm -f profiles/_variants/*.so
if [ -n "$VARIANTS" ]; then python profiles/build_variants.py $VARIANTS > /dev/null; fi
shift 0
exec /usr/local/graft/bin/gpurun "$@"
