#!/bin/bash
# 1 GPU: compute-sanitizer memcheck (one tool, as B200_PROFILING.md asks) over smoke(): every kernel family
# of the round on small inputs -- streaming, exact heads, -q1 fused, -q2/-q3 under -L, quality pipeline.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/plain_smoke.log 2>&1 &&
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_memcheck_smoke.log 2>&1
echo "memcheck rc=$?" >> $O/r02_memcheck_smoke.log
tail -12 $O/r02_memcheck_smoke.log
