import sys, os, json
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from chord_detection_b200 import ops, synth
dev=torch.device('cuda:0')
import os
FS=int(os.environ.get('FS','44100'))
N=int(FS*46.4/1000); L=(N-1)//2
NS=1_000_000 if FS==44100 else N*489
base = torch.from_numpy(np.stack([synth.s_poly(1 + i, FS, NS) for i in range(8)])).to(dev)
NC=int(os.environ.get('NC','32'))
x = base.repeat((NC+7)//8,1)[:NC].contiguous()
g=torch.Generator(device=dev).manual_seed(0)
x = x*(0.8+0.4*torch.rand(x.shape,device=dev,generator=g))
def t(fn, reps=3):
    fn(); torch.cuda.synchronize(); ts=[]
    for _ in range(reps):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts)//2]
ms=t(lambda: ops.esacf(x,FS))
r=ops.esacf(x[:2],FS,debug=True)
d=r.extra.cpu().numpy(); o=2*N+2*L
npk=d[:,o]; nfit=d[:,o+1+128]
print(json.dumps({"fs":FS,"skip":os.environ.get("CDB_ESACF_SKIP_FIT"),"frames":NC*489,"ms":ms,"peaks_per_frame_mean":float(npk.mean()),"peaks_max":float(npk.max()),"fit_mean":float(nfit.mean())}))
