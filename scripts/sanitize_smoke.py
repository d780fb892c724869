import sys, os, tempfile, numpy as np, torch
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
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
    # ragged node rows and the >255-molecule row table (edge_tc_k<..., WIDE>, linear_tc_k<..., RAGGED>)
    rng=np.random.default_rng(0)
    for B,N in ((9,20),(300,6)):
        sizes=rng.integers(1,N+1,B).astype(np.int32); sizes[0]=N
        z,t=random_batch(B,N,sizes,seed=2)
        for eng in ('strict','fast'):
            model.engine=eng
            for ragged in (False,True):
                eps=model.dynamics.forward_sizes(torch.from_numpy(t).to(dev),torch.from_numpy(z).to(dev),torch.from_numpy(sizes).to(dev),ragged=ragged)
                torch.cuda.synchronize(); print(eng,B,N,'ragged' if ragged else 'padded',float(eps.abs().max()))
    # the sampling loop (begin / step / final: tail kernel, loop state), eager and graph, padded and ragged rows
    for eng in ('strict','fp32'):
        model.engine=eng
        for graph in (False, True):
            model.use_cuda_graph=graph; model._loops={}
            torch.manual_seed(0); print(eng, 'graph' if graph else 'eager', len(model.sample(3,dev)))
    model.engine='strict'; model._loops={}
    x,h=model.sample_padded([3]*99+[40], dev); print('ragged chain', tuple(x.shape), float(x.abs().max()))
