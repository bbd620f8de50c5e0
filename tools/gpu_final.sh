#!/bin/bash
mkdir -p gpurun_out
( time python bench.py ) > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -4 gpurun_out/bench_default.err | grep real
cat gpurun_out/bench_default.json
( time python bench.py --impl reference ) > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -4 gpurun_out/bench_reference.err | grep real
cat gpurun_out/bench_reference.json
python -c "import __graft_entry__ as g; g.smoke()"
