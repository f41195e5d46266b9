#!/bin/bash
# Round-end measurement batch (one B200): tests, smoke, bench (ours + reference arm), ncu launch list, ncu --set full of the
# dominant kernels.  Everything lands in gpurun_out/ with the prefix given as $1 (default r02_final); summaries are extracted on
# the CPU side with tools/ncu_extract.py and copied into profiles/.
P=${1:-r02_final}
set -x
mkdir -p gpurun_out
timeout -s KILL 400 python -m pytest tests -x -q -m gpu > gpurun_out/${P}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${P}_pytest_gpu.log
timeout -s KILL 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${P}_smoke.log 2>&1; tail -2 gpurun_out/${P}_smoke.log
timeout -s KILL 600 python bench.py > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err; tail -c 600 gpurun_out/${P}_bench.json
timeout -s KILL 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${P}_bench_reference.json 2> gpurun_out/${P}_bench_reference.err; tail -c 400 gpurun_out/${P}_bench_reference.json
timeout -s KILL 300 python bench.py --rank 128 --steps 3 --warmup 3 --skip-svm --skip-cpu > gpurun_out/${P}_bench_rank128.json 2> gpurun_out/${P}_bench_rank128.err
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${P}_launches.csv python bench.py --steps 2 --warmup 1 --svm-rounds 2 --skip-cpu > gpurun_out/${P}_bench_under_ncu.log 2>&1
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"k_als_rows_tc64" -c 2 -f -o gpurun_out/${P}_als python bench.py --steps 1 --warmup 0 --skip-svm --skip-cpu --skip-e2e --skip-parity > gpurun_out/${P}_als_under_ncu.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:"k_svm_sweep_warp" -c 1 -f -o gpurun_out/${P}_svm python tools/svm_probe.py 10000000 28416 > gpurun_out/${P}_svm_under_ncu.log 2>&1
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:"k_als_rows_wide" -c 2 -f -o gpurun_out/${P}_wide python bench.py --rank 128 --steps 1 --warmup 0 --skip-svm --skip-cpu --skip-e2e --skip-parity > gpurun_out/${P}_wide_under_ncu.log 2>&1
(cd tools/micro && ./chol_mma_test 4 16 && ./chol_mma_test 4 16 h && ./chol_mma_test 8 4 h) > gpurun_out/${P}_solver_harness.txt 2>&1
ls -la gpurun_out | tail -14
