#!/bin/bash
# One GPU-box call that re-validates the tree as the driver will (GPU tests, smoke, default bench) and brings back
# the launch list of the config-1 training closure.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest_rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke_rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench_rc=$?"
timeout 120 python tools/probes/cfg1_closure_launches.py > gpurun_out/cfg1_closure.json 2> gpurun_out/cfg1_closure.err; echo "closure_rc=$?"; cat gpurun_out/cfg1_closure.json
timeout 120 python tools/probes/small_cfg_kernel_times.py > gpurun_out/small_cfg_kernel_times.jsonl 2> gpurun_out/small_cfg_kernel_times.err; echo "kt_rc=$?"; cat gpurun_out/small_cfg_kernel_times.jsonl
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/small_cfg_graph_launches.csv python tools/probes/small_cfg_graph_launches.py > gpurun_out/graphs_ncu.log 2>&1; echo "ncu_rc=$?"
python tools/bench_kernels.py gpurun_out/bench_n1.json
