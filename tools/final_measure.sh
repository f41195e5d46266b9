#!/bin/bash
# Round-end measurement batch (one B200): tests, smoke, bench (ours + reference arm), ncu launch list, ncu --set full of the
# dominant kernels.  Everything lands in gpurun_out/; summaries are extracted on the CPU side with tools/ncu_extract.py.
set -x
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -2 gpurun_out/pytest_gpu.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout -s KILL 600 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 600 gpurun_out/bench_full.json
timeout -s KILL 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -c 400 gpurun_out/bench_reference.json
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --svm-rounds 2 --skip-cpu > gpurun_out/bench_under_ncu.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"k_als_rows_tc64" -c 2 -o gpurun_out/final python bench.py --steps 1 --warmup 0 --skip-svm --skip-cpu --skip-e2e > gpurun_out/bench_under_ncu_full.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:"k_svm_sweep_warp" -c 1 -o gpurun_out/final_svm python tools/svm_probe.py 10000000 3552 > gpurun_out/svm_under_ncu_full.log 2>&1
ls -la gpurun_out | tail -12
