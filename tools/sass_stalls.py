import re,sys,subprocess
txt=subprocess.run(['cuobjdump','-sass',sys.argv[1]],capture_output=True,text=True).stdout
funcs=re.split(r'\n\s+Function : ',txt)[1:]
for f in funcs:
    name=f.split('\n')[0]
    lines=f.split('\n')
    n=0;tot=0;i=0;dfma=0
    while i<len(lines)-1:
        m=re.match(r'\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/',lines[i])
        if m:
            m2=re.match(r'\s+/\* (0x[0-9a-f]+) \*/',lines[i+1])
            if m2:
                hi=int(m2.group(1),16); n+=1; tot+=(hi>>41)&0xf
                if 'DFMA' in m.group(2): dfma+=1
                i+=2; continue
        i+=1
    name=subprocess.run(['c++filt',name],capture_output=True,text=True).stdout.strip()[:40]
    print(name,'instrs',n,'dfma',dfma,'sum_stall',tot,'stall/instr %.2f'%(tot/max(n,1)),'stall/dfma %.2f'%(tot/max(dfma,1)))
