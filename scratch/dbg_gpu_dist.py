import os, sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch, torch.distributed as dist
rank=int(os.environ["RANK"]); torch.cuda.set_device(rank); dev=torch.device("cuda",rank)
dist.init_process_group("nccl", device_id=dev)
from conftest import make_particles, BOX
from pylians3_b200 import dist as PD, Pk_library as PKL
from oracle import cpu as O
N=64
ctx=PD.SlabContext(N,BOX)
pos,W=make_particles(77,4*N**3,True)
mine=slice(rank,None,2)
slab=ctx.new_slab(); ctx.MA(torch.from_numpy(pos[mine].copy()).to(dev),slab,"PCS",torch.from_numpy(W[mine].copy()).to(dev))
ref=np.zeros((N,N,N),np.float32); O.MA(pos,ref,BOX,"PCS",W)
ref/=np.mean(ref,dtype=np.float64); ref-=1.0
ctx.overdensity_(slab)
x0,x1=ctx.x_range
print(rank,'delta diff rel', np.abs(slab.cpu().numpy()-ref[x0:x1]).max()/np.abs(ref).max())
dk=ctx.fft(slab).cpu().numpy(); dko=np.fft.rfftn(ref.astype(np.float64))[:,ctx.ky_range[0]:ctx.ky_range[1]]
print(rank,'fft diff', np.abs(dk-dko).max(), np.abs(dko).max())
for axis in (0,1,2):
    got=ctx.Pk(slab,axis,"PCS"); want=O.Pk(ref,BOX,axis,"PCS",1,False)
    single=PKL.Pk(ref,BOX,axis,"PCS",verbose=False) if rank==0 else None
    if rank==0:
        for nm in ("Pk","Pk1D","Pk2D","Pkphase"):
            a,b=getattr(got,nm),getattr(want,nm); c=getattr(single,nm)
            pk=np.nanmax(np.abs(b))
            print(axis,nm,'dist-vs-oracle rel',np.nanmax(np.abs(a-b)/np.abs(b)),'peakrel',np.nanmax(np.abs(a-b))/pk, '| single-vs-oracle rel',np.nanmax(np.abs(c-b)/np.abs(b)))
dist.destroy_process_group()
