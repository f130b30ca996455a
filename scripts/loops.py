import re,sys
lines=open(sys.argv[1]).read().splitlines()
ins=[]
for l in lines:
    m=re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);',l)
    if m: ins.append((int(m.group(1),16),m.group(2)))
addr={a:i for i,(a,_) in enumerate(ins)}
print("total",len(ins),"instr",ins[-1][0]/1024,"KiB")
for i,(a,t) in enumerate(ins):
    m=re.search(r'\bBRA(?:\.[A-Z.]+)?\s+(?:[!A-Z0-9, ]*?)?`?\(?\.?L?_?x?_?(\w+)\)?|BRA.*0x([0-9a-f]+)',t)
    if 'BRA' in t:
        m=re.search(r'0x([0-9a-f]+)',t)
        if m:
            tgt=int(m.group(1),16)
            if tgt<=a:
                body=ins[addr[tgt]:i+1]
                fp=sum(1 for _,x in body if re.search(r'\bD(FMA|MUL|ADD)\b',x))
                print(f"loop {tgt:#x}..{a:#x} size {(a-tgt)/1024:.1f} KiB instrs {len(body)} fp64 {fp} calls {sum(1 for _,x in body if 'CALL' in x)}")
