#!/bin/bash
for w in ztf; do
echo "== $w rounds"; PROBE_WORKLOAD=$w FULU_GP_FUSED=0 timeout 300 python scripts/fused_probe.py 20000 3 2>&1 | tail -1
echo "== $w fused (1 stepper)"; PROBE_WORKLOAD=$w FULU_GP_FUSED_STEPPERS=1 timeout 300 python scripts/fused_probe.py 20000 3 2>&1 | tail -1
echo "== $w fused (2 steppers)"; PROBE_WORKLOAD=$w FULU_GP_FUSED_STEPPERS=2 timeout 300 python scripts/fused_probe.py 20000 3 2>&1 | tail -1
echo "== $w fused serial classes"; PROBE_WORKLOAD=$w FULU_GP_FUSED_STEPPERS=1 FULU_GP_SERIAL_CLASSES=1 timeout 300 python scripts/fused_probe.py 20000 3 2>&1 | tail -1
echo "== $w rounds serial classes"; PROBE_WORKLOAD=$w FULU_GP_FUSED=0 FULU_GP_SERIAL_CLASSES=1 timeout 300 python scripts/fused_probe.py 20000 3 2>&1 | tail -1
done
echo "== plasticc rounds (inner loops unrolled, out-of-line pieces)"; FULU_GP_FUSED=0 timeout 300 python scripts/fused_probe.py 10000 3 2>&1 | tail -1
