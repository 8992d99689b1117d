"""CPU model (tests/model/uf_model.py) of how many tile-local nodes have to go to the global forest per 64x32 pixels of a
bench frame under the seam-aware BORDER rule of k_tile_build, for tiles of 64x32, 128x64 and 256x128 pixels -- i.e. what a
second-level "super-tile" merge in shared memory would leave for the global kernels.  Round-1 result: 13.2 / 5.8 / 2.6."""
import sys; import os; ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'scene-text-recognition_b200')); sys.path.insert(0,os.path.join(ROOT,'tests','model'))
import numpy as np
from ertext import synth
from oracle.refbind import PortOracle
from uf_model import build, INF, SH, MASK
port=PortOracle()
frame=synth.s_text_frame(1234)
ch=port.channels(frame)
rng=np.random.RandomState(1)
def border_count(lev,Y0,X0,TH,TW):
    tile=lev[Y0:Y0+TH,X0:X0+TW]
    par,key,find=build(tile,hi=32)
    lv=tile.ravel()
    def up(r):
        q=par[r&MASK]
        return None if q==INF else find(q)
    new=set()
    ring=[]
    for x in range(TW): ring.append((x,0,lev[Y0-1,X0+x])); ring.append((x,TH-1,lev[Y0+TH,X0+x]))
    for y in range(TH): ring.append((0,y,lev[Y0+y,X0-1])); ring.append((TW-1,y,lev[Y0+y,X0+TW]))
    for x,y,lq in ring:
        p=y*TW+x
        if lv[p]>=32 or lq>=32: continue
        a=find(key(p)); M=max(lv[p],lq)
        while True:
            u=up(a)
            if u is None or (u>>SH)>M: break
            a=u
        while a is not None and a not in new:
            new.add(a); a=up(a)
    return len(new)
res={}
for k in (0,1,4):
    lev=port.quantize(ch[k],8).reshape(ch[k].shape).astype(np.int64); lev[lev>=32]=255
    H,W=lev.shape
    for t in range(6):
        Y0=256*rng.randint(1,H//256-1) if H//256-1>1 else 256; X0=256*rng.randint(1,W//256-1)
        for (TH,TW) in ((32,64),(64,128),(128,256)):
            tot=0
            for yy in range(Y0,Y0+128,TH):
                for xx in range(X0,X0+256,TW):
                    tot+=border_count(lev,yy,xx,TH,TW)
            res.setdefault((TH,TW),[]).append(tot/16.0)   # per 64x32-equivalent area (a 256x128 block = 16 base tiles)
for k,v in res.items(): print(k,"global nodes per 64x32 of area: %.1f"%np.mean(v))
