import sys, os, tempfile, numpy as np, torch
ROOT='/root/repo'
for p in (ROOT, ROOT+'/tests', ROOT+'/tests/golden'): sys.path.insert(0,p)
from helpers import make_model, rel
dev=torch.device('cuda',0)
for name in ('forward_l2','forward_l1_pad'):
    g=np.load(f'{ROOT}/tests/golden/{name}.npz')
    with tempfile.TemporaryDirectory() as tmp:
        model=make_model(tmp,int(g['n_layers']),device=dev)
        for eng in ('fp32','strict','fast'):
            model.engine=eng
            eps=model.dynamics.forward_sizes(torch.from_numpy(g['t']).to(dev),torch.from_numpy(g['z']).to(dev),torch.from_numpy(g['sizes']).to(dev))
            print(name,eng,'rel err vs reference fixture: %.2e'%rel(eps.cpu().numpy(),g['eps']))
