import cProfile, pstats, sys, os, io
ROOT='/root/repo'
for p in (ROOT, os.path.join(ROOT,'tests','golden'), os.path.join(ROOT,'scripts')): sys.path.insert(0,p)
import torch
import bench_stage2 as B
from hierdiff_b200 import Edge_denoise
dev=torch.device('cuda',0)
m=B.load(Edge_denoise(B.VOCAB,B.F_IN,B.H,B.OUT,None,full_softmax=True),'stage2.bench.').to(dev)
batch=B.ar_batch([12,9,15,11,14],7,dev)
clone=lambda b:{k:([t.clone() for t in v] if isinstance(v,list) else v.clone()) for k,v in b.items()}
for _ in range(3): m.sample_AR(clone(batch))
torch.cuda.synchronize()
pr=cProfile.Profile(); pr.enable()
for _ in range(20): m.sample_AR(clone(batch))
torch.cuda.synchronize(); pr.disable()
s=io.StringIO(); pstats.Stats(pr,stream=s).sort_stats('tottime').print_stats(22); print(s.getvalue()[:4500])
