#!/bin/bash
# usage: sass_local.sh <object.o> <kernel-name-substring> : local-memory (spill) instructions of a kernel by source line
set -e
T=$(mktemp -d); cd $T; cuobjdump -xelf all "$1" >/dev/null; nvdisasm -g -c *.cubin > all.sass 2>/dev/null
python - "$2" <<'PY'
import re,sys
from collections import Counter
cur=None; fn=None; cnt=Counter(); tot=Counter()
for line in open('all.sass'):
    m=re.search(r'//## File "([^"]+)", line (\d+)',line)
    if m: cur=(m.group(1).split('/')[-1],int(m.group(2))); continue
    m=re.match(r'\s*\.text\.(\S+):',line)
    if m: fn=m.group(1)
    if fn and sys.argv[1] in fn:
        if re.search(r'^\s+/\*[0-9a-f]{4}\*/',line): tot[cur]+=1
        if re.search(r'\b(STL|LDL)\b',line): cnt[(cur,'STL' if 'STL' in line else 'LDL')]+=1
for k,v in sorted(cnt.items(), key=lambda x:-x[1])[:25]: print(k,v)
print('total instructions', sum(tot.values()))
PY
rm -rf $T
