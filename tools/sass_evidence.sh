#!/bin/bash
# Blackwell evidence for the benchmarked kernel variants (runs here, no GPU): SASS mnemonic counts that prove
# the TMA-engine bulk copy (UBLKCP), mbarrier (SYNCS), elected thread (ELECT), 16-byte streaming stores
# (STG.E.EF.128), programmatic dependent launch (ACQBULK / PREEXIT), plus registers / spills / shared memory.
#   tools/sass_evidence.sh > profiles/r02_sass_evidence.txt
LIB=pogema_b200/_lib/libpgm_b200.so
declare -A K=(
  ["1 configs[1] / configs[4] share r=5 : fast<TEAM 64, APT 1, priority, r 5>"]="_ZN3pgm20pgm_fast_step_kernelILi64ELi1ELi0ELi5EEEvNS_8StepArgsE"
  ["2 configs[4] r=3 share : fast<64, 1, priority, r 3>"]="_ZN3pgm20pgm_fast_step_kernelILi64ELi1ELi0ELi3EEEvNS_8StepArgsE"
  ["3 configs[4] on one GPU (16384 instances, two agents per thread), r=3 : fast<32, 2, priority, r 3>"]="_ZN3pgm20pgm_fast_step_kernelILi32ELi2ELi0ELi3EEEvNS_8StepArgsE"
  ["4 configs[4] r=7 share : fast<64, 1, priority, r 7>"]="_ZN3pgm20pgm_fast_step_kernelILi64ELi1ELi0ELi7EEEvNS_8StepArgsE"
  ["5 configs[2] : fast<256, 1, soft, r 5>"]="_ZN3pgm20pgm_fast_step_kernelILi256ELi1ELi2ELi5EEEvNS_8StepArgsE"
  ["6 configs[3] : fast<256, 4, block_both, r 5>"]="_ZN3pgm20pgm_fast_step_kernelILi256ELi4ELi1ELi5EEEvNS_8StepArgsE"
  ["7 generic kernel (reset / observe / uncommon shapes) : step<32, priority, r 5, dense grid>"]="_ZN3pgm15pgm_step_kernelILi32ELi0ELi0ELi5ELi0ELi0EEEvNS_8StepArgsE"
)
echo "cuobjdump -sass / -res-usage of $LIB ($(date -u +%F)), nvcc $(nvcc --version | grep -o 'release [0-9.]*')"
echo "arch list: $(cuobjdump -lelf $LIB | grep -o 'sm_[0-9a-z]*' | sort -u | tr '\n' ' ')"
for name in $(printf "%s\n" "${!K[@]}" | sort | tr " " "~"); do
  name=${name//\~/ }
  f=${K[$name]}
  echo
  echo "== $name"
  echo "   $f"
  cuobjdump -res-usage -fun "$f" $LIB 2>/dev/null | grep -E "REG|SHARED" | head -1 | sed 's/^/   /'
  cuobjdump -sass -fun "$f" $LIB 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed 's/^\s*\/\*[0-9a-f]*\*\/\s*//; s/\s*\/\*.*$//' > /tmp/_k.sass
  echo "   SASS instructions: $(wc -l < /tmp/_k.sass)"
  for m in UBLKCP SYNCS ELECT "STG.E.EF.128" "STG.E.EF" ACQBULK PREEXIT "LDS.64" "LDS.U16" ATOMS "BAR.SYNC" "BAR.RED" WARPSYNC SHFL REDUX STL LDL "NANOSLEEP"; do
    c=$(grep -c -- "$m" /tmp/_k.sass)
    [ "$c" != "0" ] && printf "   %-14s %s\n" "$m" "$c"
  done
  echo "   excerpt (bulk copy + mbarrier):"
  grep -n -E "UBLKCP|SYNCS|ELECT|ACQBULK|PREEXIT" /tmp/_k.sass | head -8 | sed 's/^/      /'
done
