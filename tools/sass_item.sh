#!/bin/bash
# Quick SASS iteration on the headline instantiation of the item kernel: compiles regex_item.cu with only
# k_chain_item<4,1,SPEC> (~6 s), prints registers / spills and static instruction counts, leaves /tmp/ki.sass + /tmp/ki.dis.
#   [SPEC=3] tools/sass_item.sh [extra nvcc flags]
set -e
HERE="$(cd "$(dirname "$0")/.." && pwd)"
OUT="${TMPDIR:-/tmp}"
cd "$HERE/custrings_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -I ../../include -Xptxas -v \
     -DITEM_EXPERIMENT=${SPEC:-3} "$@" -cubin regex_item.cu -o "$OUT/ki.cubin" 2> "$OUT/ki.ptxas"
grep -A2 "k_chain_itemILi4ELi1ELi" "$OUT/ki.ptxas" | grep -E "registers|spill" | head -3
cuobjdump -sass "$OUT/ki.cubin" | awk '/Function : .*k_chain_itemILi4ELi1ELi/{f=1} f' | grep -E "^\s+/\*[0-9a-f]{4}\*/" > "$OUT/ki.sass"
nvdisasm -gi -c "$OUT/ki.cubin" > "$OUT/ki.dis"
echo "total $(wc -l < "$OUT/ki.sass") BRA.DIV $(grep -c 'BRA.DIV' "$OUT/ki.sass") BRA $(grep -c 'BRA' "$OUT/ki.sass") ISETP $(grep -c ISETP "$OUT/ki.sass") LDL/STL $(grep -c 'LDL\|STL' "$OUT/ki.sass")"
