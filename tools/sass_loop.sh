#!/bin/bash
# Static size of the 4-base inner loops of pileup_count_kernel<true> (SASS instructions between the
# 16-bit base load and the loop's back edge).  Usage: tools/sass_loop.sh
cd "$(dirname "$0")/../longsom_b200/csrc" || exit 1
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo --extended-lambda -cubin -o /tmp/ls_pileup.cubin ls_pileup.cu 2>/dev/null || exit 1
cuobjdump -sass /tmp/ls_pileup.cubin | awk '/Function : .*pileup_count_kernelILb1/{f=1;next} /Function : /{f=0} f' | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | sed 's/ *\/\* 0x[0-9a-f]* \*\/$//' > /tmp/sass_k1.txt
python3 - <<'PY'
import re
L=[l.rstrip() for l in open('/tmp/sass_k1.txt')]
addr=lambda l:int(re.search(r'/\*([0-9a-f]+)\*/',l).group(1),16)
idx=[i for i,l in enumerate(L) if 'LDG.E.U16' in l]
print("kernel SASS lines:",len(L))
for i in idx:
    # find the next backward branch that jumps to an address <= this load's address
    a0=addr(L[i])
    for j in range(i+1,len(L)):
        m=re.search(r'BRA\s+(?:U,\s*)?(0x[0-9a-f]+)',L[j])
        if m and int(m.group(1),16)<=a0 and int(m.group(1),16)>a0-0x400:
            tgt=int(m.group(1),16)
            start=[k for k in range(i,-1,-1) if addr(L[k])==tgt]
            s=start[0] if start else i
            body=L[s:j+1]
            n_atom=sum('ATOMS' in x or 'RED' in x for x in body)
            print("loop @%#x..%#x: %d instr, %d shared atomics/reds"%(tgt,addr(L[j]),len(body),n_atom)); break
PY
