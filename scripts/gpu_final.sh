#!/bin/bash
# round 2, final build: GPU suite, smoke, the three workloads' bench lines with their CPU baselines, the reference arm, one ncu --set full
# capture of the headline evaluation kernel (after the same command ran clean without ncu) and the launch list of one headline step
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/z_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/z_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/z_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/z_smoke.log
timeout 900 python bench.py > gpurun_out/z_bench.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/z_bench.log | cut -c1-200
timeout 900 python bench.py --impl reference > gpurun_out/z_bench_ref.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/z_bench_ref.log | cut -c1-200
timeout 900 python bench.py --workload ztf > gpurun_out/z_bench_ztf.log 2>&1; echo "ztf rc=$?"; tail -1 gpurun_out/z_bench_ztf.log | cut -c1-200
timeout 1200 python bench.py --workload ddf > gpurun_out/z_bench_ddf.log 2>&1; echo "ddf rc=$?"; tail -1 gpurun_out/z_bench_ddf.log | cut -c1-200
CMD="python bench.py --steps 1 --warmup 1 --curves 2000 --no-cpu-baseline --no-library-baseline"
timeout 600 $CMD > gpurun_out/z_plain_bench.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:fit_eval_kernel -s 10 -c 2 -f -o gpurun_out/prof_eval_final $CMD > gpurun_out/z_ncu_eval.log 2>&1; echo "ncu eval rc=$?"
python scripts/ncu_digest.py gpurun_out/prof_eval_final.ncu-rep gpurun_out/r02_eval_kernel_ncu_final.json
timeout 600 $CMD > gpurun_out/z_plain_bench2.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 800 --csv --log-file gpurun_out/r02_launches_final.csv $CMD > gpurun_out/z_ncu_list.log 2>&1; echo "launch list rc=$?"
