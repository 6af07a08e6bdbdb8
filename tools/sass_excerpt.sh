#!/bin/bash
# tools/sass_excerpt.sh [lib.so] > profiles/rN_sass_fast6_excerpt.txt : SASS evidence of the Blackwell-native path
# (runs here: cuobjdump needs no GPU).
LIB=${1:-vk_compute_mipmaps_b200/libnvpyr.so}
K=${2:-fastSrgba8KernelILi6ELb0ELb0ELb0ELi24EEE}
echo "# SASS evidence, $LIB built by __graft_entry__.build() (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a)"
echo "# made by: tools/sass_excerpt.sh (tools/sass_loop.sh + cuobjdump -sass | grep)"
echo; echo "== architectures in the library"
cuobjdump -lelf $LIB | sort | uniq -c
echo; echo "== $K: loop structure and opcode histogram (tools/sass_loop.sh; the slab loop is the ~410-instruction one)"
bash tools/sass_loop.sh $LIB $K 2>/dev/null > /tmp/k_loops.txt; head -8 /tmp/k_loops.txt | cut -c1-230
cuobjdump -sass $LIB | awk -v pat="$K" '/Function :/{f=($0 ~ pat)} f' > /tmp/k_full.sass
echo; echo "== Blackwell-specific mnemonics in that kernel (count)"
grep -oE "UTMALDG[.A-Z0-9]*|SYNCS[.A-Z0-9]*|FADD2|FFMA2|FMUL2|ELECT|ACQBULK|UBLKCP|LDGSTS[.A-Z0-9]*" /tmp/k_full.sass | sort | uniq -c
echo; echo "== the encode look-up of the slab loop: FFMA (z = S' * 2^k + c), PRMT (key(z) << 8 | lane << 2), LDS from the lane-private row table, IADD3"
grep -E "FFMA R[0-9]+, R[0-9]+, R[0-9]+, 0\.03118|PRMT R[0-9]+, R[0-9]+, 0x5324|LDS R[0-9]+, \[R[0-9]+\+UR[0-9]+\+-0x3cfe80\]|IADD3 R[0-9]+, PT, PT, R[0-9]+, 0x31000000" /tmp/k_full.sass | head -12
echo; echo "== the TMA issue + mbarrier wait of the slab loop (every line with UTMALDG / SYNCS and its neighbours)"
grep -E -B2 -A2 "UTMALDG|SYNCS" /tmp/k_full.sass | sed -E 's/\s+\/\* 0x[0-9a-f]+ \*\///' | head -70
echo; echo "== library-wide counts"
cuobjdump -sass $LIB | grep -oE "UTMALDG|SYNCS|FADD2|FFMA2|FMUL2|LDGSTS" | sort | uniq -c
