#!/bin/bash
# The block code under ASan+UBSan and TSan on the host-thread emulation, full matrix (about 5.5 minutes on 8 cores);
# tests/test_sanitize_host.py runs a slice of it and explains what each sanitizer stands in for.  No GPU needed.
set -e
cd "$(dirname "$0")/../tests/_host"
mkdir -p ../_build
g++ -std=c++20 -O1 -g -pthread -fsanitize=address,undefined -fno-omit-frame-pointer sanitize_main.cpp -o ../_build/sanitize_asan
g++ -std=c++20 -O1 -g -pthread -fsanitize=thread sanitize_main.cpp -o ../_build/sanitize_tsan
g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer sanitize_lbfgsb_main.cpp -o ../_build/sanitize_lbfgsb
cd ../_build
./sanitize_lbfgsb
for n in 1 2 7 8 30 100 103 104 160; do for nthr in 32 128 256 512; do for fam in 0 32 34 64; do
  ks=$(( fam == 0 ? 1 : 0 ))
  for exe in sanitize_asan sanitize_tsan; do
    out=$(./$exe $n $nthr $fam $ks 8 2>&1) || { echo "FAIL $exe n=$n nthr=$nthr family=$fam"; echo "$out" | head -40; exit 1; }
    if echo "$out" | grep -q "Sanitizer\|runtime error"; then echo "FAIL $exe n=$n nthr=$nthr family=$fam"; echo "$out" | head -40; exit 1; fi
  done
done; done; done
echo "clean: 9 sizes x 4 block sizes x 4 families x {asan+ubsan, tsan}"
