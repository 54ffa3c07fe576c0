"""Static view of a kernel's SASS schedule: decodes the per-instruction stall count from the control bits of the 128-bit
encoding (bits 105-108) in `cuobjdump -sass` output and sums it per opcode class.  The sum is the issue time of ONE warp
running the listing once with no contention — a lower bound that lets two builds of the same kernel be compared without
a GPU.  usage: cuobjdump -sass lib.so | awk '/Function : .*KERNEL/{f=1} /Function : /{if(!/KERNEL/)f=0} f' > k.sass;
python tools/sass_stalls.py k.sass [other.sass ...]"""
import re,sys,collections
def analyse(path):
    lines=open(path).read().splitlines()
    ins=[]
    i=0
    pat=re.compile(r'^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* (0x[0-9a-f]{16}) \*/')
    pat2=re.compile(r'^\s+/\* (0x[0-9a-f]{16}) \*/')
    while i<len(lines):
        m=pat.match(lines[i])
        if m and i+1<len(lines):
            m2=pat2.match(lines[i+1])
            if m2:
                hi=int(m2.group(1),16)
                stall=(hi>>41)&0xf; yld=(hi>>45)&1; wr=(hi>>46)&7; rd=(hi>>49)&7; wait=(hi>>52)&0x3f
                op=m.group(2).split()[0] if not m.group(2).startswith('@') else m.group(2).split()[1]
                ins.append((op,stall,yld,wr,rd,wait,m.group(2)))
                i+=2; continue
        i+=1
    tot=sum(s for _,s,*_ in ins)
    by=collections.Counter(); cnt=collections.Counter()
    for op,s,*_ in ins:
        k=op.split('.')[0]; by[k]+=s; cnt[k]+=1
    print(path, 'instructions', len(ins), 'sum of stall counts', tot, 'avg', round(tot/len(ins),2))
    for k,v in by.most_common(8): print('   ',k, 'n',cnt[k],'stall sum',v,'avg',round(v/cnt[k],2))
    hist=collections.Counter(s for _,s,*_ in ins); print('    stall histogram', sorted(hist.items()))
    return ins
for p in sys.argv[1:]:
    analyse(p)
