#!/bin/bash
# fused fit (one persistent CTA per SM, evaluation groups + stepper warp): parity suite, then A/B against the rounds path
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/f1_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/f1_pytest_gpu.log
B="python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-library-baseline"
for cfg in "FULU_GP_FUSED=0" "FULU_GP_FUSED=1" "FULU_GP_FUSED_SLOTS=8" "FULU_GP_FUSED_SLOTS=12" "FULU_GP_FUSED_SLOTS=24" "FULU_GP_FUSED_SLOTS=32" "FULU_GP_FUSED_ROTATE=0"; do
  echo "== $cfg"
  env $cfg timeout 300 $B 2>&1 | tail -1 | python -c "
import sys, json
try:
    j = json.loads(sys.stdin.read())
    r = j['roofline']
    print('value %.0f e2e %.0f ms/step %.2f eval_ms %.2f step_ms %.2f rounds %.0f frac %.4f evals/curve %.3f flagged %d mean_lml %.9f' % (j['value'], j['e2e']['value'], j['ms_per_step'], r['eval_kernel_ms_per_step'], r['step_kernel_ms_per_step'], r['rounds_per_step'], r['frac'], r['evals_per_curve'], j['parity']['curves_flagged'], j['parity']['mean_lml']))
except Exception as e:
    print('bench failed', e)
"
done
