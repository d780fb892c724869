import sys, os, tempfile, numpy as np, torch
ROOT='/root/repo'
for p in (ROOT, ROOT+'/tests', ROOT+'/tests/golden'): sys.path.insert(0,p)
from helpers import make_model, rel
dev=torch.device('cuda',0)
with tempfile.TemporaryDirectory() as tmp:
    model=make_model(tmp,4,timesteps=1000,device=dev,engine='fp32')
    sizes=[40]*64
    out={}
    for eng in ('fp32','strict','fast'):
        model.engine=eng
        torch.manual_seed(0)
        x,h=model.sample_padded(sizes,dev)
        out[eng]=(x.numpy(),h.numpy())
        print(eng, np.abs(x.numpy()).max(), np.abs(h.numpy()).max())
    for eng in ('strict','fast'):
        print(eng,'x rel',rel(out[eng][0],out['fp32'][0]),'h rel',rel(out[eng][1],out['fp32'][1]))
