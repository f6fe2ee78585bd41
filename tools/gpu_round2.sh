#!/bin/bash
# GPU-box session for the quality pipeline: smoke, its tests first (fail fast), the whole GPU suite,
# qbench lines, a memcheck pass over the small quality tests -> gpurun_out/
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -4 gpurun_out/smoke.log
timeout 600 python -m pytest tests/test_quality_gpu.py -m gpu -q -s 2>&1 | tail -40 > gpurun_out/pytest_quality.log
tail -25 gpurun_out/pytest_quality.log
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for c in quality_cfg3 quality_cfg1 quality_cfg2; do
  timeout 300 python tools/qbench.py --config $c > gpurun_out/qbench_$c.json 2> gpurun_out/qbench_$c.err || tail -5 gpurun_out/qbench_$c.err
  cat gpurun_out/qbench_$c.json
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err; cut -c1-400 gpurun_out/bench_cfg2.json
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_quality_gpu.py -m gpu -q -k "fixture or stream" > gpurun_out/memcheck_quality.log 2>&1; tail -5 gpurun_out/memcheck_quality.log
