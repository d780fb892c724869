import sys, os, tempfile, numpy as np, torch
ROOT='/root/repo'
for p in (ROOT, ROOT+'/tests', ROOT+'/tests/golden'): sys.path.insert(0,p)
from helpers import make_model, random_batch
dev=torch.device('cuda',0)
with tempfile.TemporaryDirectory() as tmp:
    model=make_model(tmp,1,timesteps=3,device=dev,engine='strict')
    for sizes,N in (([5,3,8],8),([40,17],40)):
        z,t=random_batch(len(sizes),N,sizes,seed=1)
        for eng in ('strict','fast'):
            model.engine=eng
            eps=model.dynamics.forward_sizes(torch.from_numpy(t).to(dev),torch.from_numpy(z).to(dev),torch.tensor(sizes,dtype=torch.int32,device=dev))
            torch.cuda.synchronize(); print(eng,sizes,float(eps.abs().max()))
    torch.manual_seed(0); print(len(model.sample(3,dev)))
