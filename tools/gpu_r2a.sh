#!/bin/bash
# round 2, call A: parity suite, smoke, bench (short stream leg), reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2a_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2a_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 --stream-seconds 12 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; tail -c 6000 gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2a_ref.json 2> gpurun_out/r2a_ref.err; echo "ref rc=$?"; tail -c 1500 gpurun_out/r2a_ref.json
