import re,csv,sys,subprocess
from collections import defaultdict
rep, obj, kern = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv)>4 else 40
import os; subprocess.run("rm -f /tmp/*.cubin; cd /tmp; cuobjdump -xelf all %s >/dev/null 2>&1" % os.path.abspath(obj), shell=True)
import glob
cub=glob.glob('/tmp/*.cubin')[0]
dis=subprocess.run("nvdisasm -g %s"%cub, shell=True, capture_output=True, text=True).stdout.split('\n')
srccsv=subprocess.run("ncu -i %s --page source --csv"%rep, shell=True, capture_output=True, text=True).stdout
rows=list(csv.reader(srccsv.split('\n')))
kname=rows[0][1]
print(kname[:150])
# find section for the mangled name containing kern and matching the instruction count
h=rows[1]; ie=h.index('Instructions Executed'); ia=h.index('Address'); ist=h.index('# Samples'); isrc=h.index('Source')
data=[r for r in rows[2:] if len(r)>ie]
ninstr=len(data)
secs=[i for i,l in enumerate(dis) if l.startswith('.text.') and kern in l]
best=None
for st in secs:
    a2l={}; cur=None
    for l in dis[st+1:]:
        if l.startswith('.text.') or l.startswith('.section'): break
        m=re.search(r'//## File "([^"]+)", line (\d+)',l)
        if m: cur=(m.group(1).split('/')[-1],int(m.group(2))); continue
        m=re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);',l)
        if m: a2l[int(m.group(1),16)]=(cur,m.group(2))
    if len(a2l)==ninstr: best=a2l; break
if best is None: print("no matching section", ninstr, [len(x) for x in []]); sys.exit(1)
base=int(data[0][ia],16)
agg=defaultdict(lambda:[0,0]); tot=0; tots=0
for r in data:
    a=int(r[ia],16)-base; n=int(r[ie]); s=int(r[ist])
    key=best.get(a,(None,''))[0]
    agg[key][0]+=n; agg[key][1]+=s; tot+=n; tots+=s
print("total warp instr", tot, "samples", tots)
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][0])[:top]:
    print("%10d %5.1f%%  samples %5.1f%%  %s"%(v[0],100*v[0]/tot,100*v[1]/max(tots,1),k))
