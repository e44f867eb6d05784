#!/bin/bash
# round 2 evidence: ncu --set full of the tiled long-curve evaluation (N = 1024, 2048) and of the headline evaluation kernel, the launch
# list of one headline step.  Every ncu run follows the same command run clean without ncu.  The reports are digested here
# (scripts/ncu_digest.py) and deleted: with --import-source they exceed what gpurun brings back.
mkdir -p gpurun_out
for n in 1024 2048; do
  CMD="python scripts/long_curve_probe.py $n"
  timeout 300 $CMD > gpurun_out/g_plain_$n.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:lml_kernel -s 1 -c 1 -f -o /tmp/prof_r02_tiled$n $CMD > gpurun_out/g_ncu_$n.log 2>&1; echo "ncu tiled $n rc=$?"; tail -1 gpurun_out/g_plain_$n.log
  python scripts/ncu_digest.py /tmp/prof_r02_tiled$n.ncu-rep gpurun_out/r02_tiled${n}_ncu.json
done
CMD="python bench.py --steps 1 --warmup 1 --curves 2000 --no-cpu-baseline --no-library-baseline"
timeout 600 $CMD > gpurun_out/g_plain_bench.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:fit_eval_kernel -s 10 -c 2 -f -o /tmp/prof_r02_eval $CMD > gpurun_out/g_ncu_eval.log 2>&1; echo "ncu eval rc=$?"
python scripts/ncu_digest.py /tmp/prof_r02_eval.ncu-rep gpurun_out/r02_eval_kernel_ncu.json
timeout 600 $CMD > gpurun_out/g_plain_bench2.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 800 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/g_ncu_list.log 2>&1; echo "launch list rc=$?"
ls -la gpurun_out | head -30; du -sh gpurun_out
