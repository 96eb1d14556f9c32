#!/usr/bin/env bash
# usage: tools/f32_loop_probe.sh <variant> <unroll>   -> prints loop length and .reuse count of the fp32 hot loops
set -euo pipefail
cd "$(dirname "${BASH_SOURCE[0]}")/.."
v=$1; u=$2; pb=${3:-0}; o=/tmp/f32probe_${v}_${u}_${pb}
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -fmad=false -DNB_F32_ACC_VARIANT=$v -DNB_F32_UNROLL=$u -DNB_F32_PERTURB=$pb -Xptxas -v -c tools/f32_loop_probe.cu -o $o.o 2> $o.log
cuobjdump -sass $o.o > $o.sass
python3 - $o.sass "$v/p$pb" $u <<'PY'
import re,sys
t=open(sys.argv[1]).read()
for f in re.split(r'\n\s*Function : ', t)[1:]:
    name=f.split('\n',1)[0].strip()
    lines=[re.sub(r'\s*/\* 0x[0-9a-f]+ \*/','',l).rstrip() for l in f.split('\n') if re.search(r'/\*[0-9a-f]{4}\*/', l)]
    addr=lambda l:int(re.search(r'/\*([0-9a-f]{4})\*/',l).group(1),16)
    best=None
    for i,l in enumerate(lines):
        m=re.search(r'BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?(0x[0-9a-f]+)',l)
        if m and int(m.group(1),16)<addr(l):
            j=[k for k,x in enumerate(lines) if addr(x)==int(m.group(1),16)]
            if j:
                body=lines[j[0]:i+1]
                if sum('MUFU' in x for x in body)>=8 and (best is None or len(body)<len(best)): best=body
    if best is None: continue
    nm=sum('MUFU' in x for x in best)
    print(f"variant {sys.argv[2]} unroll {sys.argv[3]} {name[22:62]:40s} loop {len(best):4d} instr / {nm} MUFU = {len(best)/(nm/2):.2f} per group, reuse {sum(l.count('.reuse') for l in best)} ({sum(l.count('.reuse') for l in best)/(nm/2):.2f}/group)")
PY
grep -E "registers|spill" $o.log | grep -v "0 bytes spill" | sed 's/ptxas info    ://' | tr '\n' ' '; echo
