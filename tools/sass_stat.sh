#!/bin/bash
# usage: sass_stat.sh  -> compiles regex_bits.cu to cubin and prints stats for k_chain64<4,1>
set -e
cd /root/repo/custrings_b200/csrc
time nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --expt-relaxed-constexpr -I ../../include -Xptxas -v -DCUSTR_EXPERIMENT_ONLY_4_1 -cubin regex_bits.cu -o /tmp/rb.cubin 2> /tmp/rb.ptxas
grep -A1 "k_chain64ILi4ELi1E" /tmp/rb.ptxas | grep -E "registers|spill" | head -3
cuobjdump -sass -fun '_ZN5custr4bits9k_chain64ILi4ELi1EEEvNS0_8ChainDevENS0_4ArgsE' /tmp/rb.cubin | grep -E "^\s+/\*[0-9a-f]{4}\*/" > /tmp/k.sass
echo "total $(wc -l < /tmp/k.sass) BRA.DIV $(grep -c 'BRA.DIV' /tmp/k.sass) WARPSYNC $(grep -c WARPSYNC /tmp/k.sass) BRA $(grep -c 'BRA' /tmp/k.sass) ISETP $(grep -c ISETP /tmp/k.sass) S2R $(grep -c 'S2R\|S2UR' /tmp/k.sass) LDL $(grep -c 'LDL\|STL' /tmp/k.sass)"
