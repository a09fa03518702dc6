import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, amv_codec_tools_b200 as amv
from oracle_lib import *
ctx=amv.AmvCuda(0); o=Oracle()
w,h=48,40
y,u,v=synth_frames(11,w,h,seed=61,kind='flat')
pk,off,sz=o.encode_frames(y,u,v,w,h,2)
print(sz)
wy,wu,wv,wst=o.decode_frames(pk,off,sz,w,h)
for lp in range(6):
    ctx.set_option("decode_log2_lanes",lp)
    dy,du,dv,st=ctx.decode_frames(pk,off,sz,w,h)
    print(lp, st, [int((dy[i]!=wy[i]).sum()) for i in range(11)], ctx.get_stat("decode_sync_rounds"))
