import re,sys,collections
lines=open(sys.argv[1]).read().splitlines()
lo=int(sys.argv[2],16); hi=int(sys.argv[3],16)
c=collections.Counter()
for l in lines:
    m=re.match(r'\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);',l)
    if m:
        a=int(m.group(1),16)
        if lo<=a<=hi:
            t=[o for o in m.group(2).split() if not o.startswith('@')]
            c[t[0].split('.')[0] if t[0]!='IMAD.MOV.U32' else 'IMAD.MOV']+=1
            if t[0].startswith('IMAD.MOV'): c['(mov)']+=1
print(sum(c.values())-c['(mov)'], c.most_common(40))
