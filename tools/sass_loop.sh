#!/bin/bash
# tools/sass_loop.sh <lib.so> [mangled-substring]: SASS of one kernel with the loop structure (backward branches)
# and an opcode histogram of its innermost hot loop (the longest backward-branch span is printed first).
LIB=$1; PAT=${2:-fastSrgba8KernelILi6ELb0ELb0ELb0ELi24EEE}
cuobjdump -sass $LIB | awk -v pat="$PAT" '/Function :/{f=($0 ~ pat)} f' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s*\/\*([0-9a-f]+)\*\/\s+/\1 /; s/\s*\/\*.*//; s/ +/ /g' > /tmp/kernel.sass
echo "instructions: $(wc -l < /tmp/kernel.sass)"
python3 - <<'PY'
import re,collections
L=[l.split(' ',1) for l in open('/tmp/kernel.sass').read().splitlines()]
addr=[int(a,16) for a,_ in L]
loops=[]
for a,(x,ins) in zip(addr,L):
    m=re.search(r'BRA(?:\.[A-Z.]+)? (?:[!U]*P\d, )?0x([0-9a-f]+)',ins)
    if m and 'BRA.DIV' not in ins:
        t=int(m.group(1),16)
        if t<a: loops.append((t,a))
for t,a in sorted(loops,key=lambda z:z[0]-z[1]):
    body=[ins for ad,(x,ins) in zip(addr,L) if t<=ad<=a]
    ops=collections.Counter()
    for ins in body:
        ins=re.sub(r'^@!?U?P\d+ ','',ins)
        op=ins.split()[0]
        op='MOVE' if op in('MOV','IMAD.MOV.U32','CS2R') else op.split('.')[0]
        ops[op]+=1
    print(f"loop {t:#x}..{a:#x}: {len(body)} instr:", ' '.join(f"{k}:{v}" for k,v in ops.most_common(14)))
PY
