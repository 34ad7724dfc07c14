#!/bin/bash
# Runs the product's device code under the CPU emulator (tests/emu) built with AddressSanitizer +
# UBSan: catches out-of-bounds accesses beyond the shared-memory / global arenas and undefined shifts
# in the kernels' index arithmetic. Test infrastructure only. Usage: bash scripts/emu_asan.sh
set -e
cd "$(dirname "$0")/.."
keep=$(mktemp)
[ -f tests/emu/libafq_emu.so ] && cp tests/emu/libafq_emu.so "$keep"
g++ -O1 -g -std=c++17 -fPIC -ffp-contract=off -DAFQ_EMU -fsanitize=address,undefined -fno-omit-frame-pointer \
    -Itests/emu -shared -o tests/emu/libafq_emu.so tests/emu/emu_pipeline.cpp
touch tests/emu/libafq_emu.so
export LD_PRELOAD=$(gcc -print-file-name=libasan.so)
export ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0:halt_on_error=1
for lim in 0 2500 6000; do
  echo "== AFQ_PS_LIMIT_WORDS=$lim"
  if [ $lim = 0 ]; then python scripts/emu_asan_cases.py; else AFQ_PS_LIMIT_WORDS=$lim python scripts/emu_asan_cases.py; fi
done
unset LD_PRELOAD
[ -s "$keep" ] && cp "$keep" tests/emu/libafq_emu.so && touch tests/emu/libafq_emu.so
rm -f "$keep"
