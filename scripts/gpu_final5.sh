#!/bin/bash
B="python bench.py --kernel rbf_matern --no-library-baseline --no-cpu-baseline --steps 2 --warmup 2"
echo "== lean paths (P <= 5)"; timeout 300 $B 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'])"
echo "== compact paths"; FULU_GP_LIB=$PWD/fulu_b200/_lib/libfulu_gp_l4.so timeout 300 $B 2>&1 | tail -1 | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['value'], j['ms_per_step'])"
