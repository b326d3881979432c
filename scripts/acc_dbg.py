import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch, bench
from gyeeta_b200 import engine as ge
dev=torch.device('cuda',0); n=100_000_000
evs=[bench.gen_events_gpu(torch,n,1234+7919*b,0,1,dev) for b in range(2)]
eng=ge.Engine(device=0,max_svcs=1<<17,max_tasks=1<<15,max_batch=(1<<27)-1,stage_batch=1<<23)
for r in range(6):
    eng.ingest_device_ptr(evs[r%2].data_ptr(), n)
eng.sync()
w0=evs[0][:,0]; isr=((evs[0][:,3]>>32)&0xFFFF)==5
u,c=torch.unique(w0[:2_000_000][isr[:2_000_000]],return_counts=True)
o=torch.argsort(c,descending=True)
for k in (0,1,40,41,400,401,402,403):
    sid=int(u[o[k]])
    vals=torch.cat([(e[:,2][(e[:,0]==sid)&(((e[:,3]>>32)&0xFFFF)==5)]&0xFFFFFFFF) for e in evs]).double()
    ex=torch.quantile(vals[:16_000_000], torch.tensor([.5,.95,.99],device=dev,dtype=torch.float64), interpolation='lower').cpu().numpy()
    got=eng.quantiles(sid&0xFFFFFFFFFFFFFFFF,[.5,.95,.99])
    m,w,mn,mx=eng.export_tdigest(sid&0xFFFFFFFFFFFFFFFF)
    print(k, vals.numel(), np.round((got-ex)/ex*100,3), len(m), int(w.sum()), mn, mx, float(vals.min()), float(vals.max()))
