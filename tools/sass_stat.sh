#!/bin/bash
# Quick SASS iteration on the headline instantiation: compiles regex_bits.cu with only k_chain64<4,1> instantiated
# (-DCUSTR_EXPERIMENT_ONLY_4_1, ~3 s) and prints registers / spills and a few static instruction counts.
#   [SPEC=3] tools/sass_stat.sh [extra nvcc flags]      SPEC = shape specialisation of regex_chain64.cuh (0 = generic)
set -e
HERE="$(cd "$(dirname "$0")/.." && pwd)"
OUT="${TMPDIR:-/tmp}"
cd "$HERE/custrings_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -I ../../include -Xptxas -v \
     -DCUSTR_EXPERIMENT_ONLY_4_1=${SPEC:-0} "$@" -cubin regex_bits.cu -o "$OUT/rb.cubin" 2> "$OUT/rb.ptxas"
grep -A2 "k_chain64ILi4ELi1ELi" "$OUT/rb.ptxas" | grep -E "registers|spill" | head -3
cuobjdump -sass -fun "_ZN5custr4bits9k_chain64ILi4ELi1ELi${SPEC:-0}EEEvNS0_8ChainDevENS0_4ArgsE" "$OUT/rb.cubin" | grep -E "^\s+/\*[0-9a-f]{4}\*/" > "$OUT/k.sass"
echo "total $(wc -l < "$OUT/k.sass") BRA.DIV $(grep -c 'BRA.DIV' "$OUT/k.sass") BRA $(grep -c 'BRA' "$OUT/k.sass") ISETP $(grep -c ISETP "$OUT/k.sass") S2R $(grep -c 'S2R\|S2UR' "$OUT/k.sass") LDL/STL $(grep -c 'LDL\|STL' "$OUT/k.sass")"
