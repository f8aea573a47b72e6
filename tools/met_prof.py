import sys, time, json
sys.path.insert(0,'/root/repo')
import numpy as np, torch
import rabbit_transcoding_b200 as rb
kw=dict(bitdepth=10, width=1280, scale=0.68, height_blocks=80, seed=0x0AB817, transfer_filter=0)
gof=rb.synthetic.generate_gof_parallel(32, workers=16, **kw)
codec=rb.codec.PCCCodecB200(device=0)
codec.uploadGof(gof); codec.decodeGof()
mp=rb.metrics.default_parameters(resolution=1023.0)
met=rb.metrics.PCCMetricsB200(codec); met.setParameters(mp)
srcs=[dict(positions=torch.from_numpy(s["positions"]).pin_memory().numpy(), colors=torch.from_numpy(s["colors"]).pin_memory().numpy(), normals=torch.from_numpy(s["normals"]).pin_memory().numpy()) for s in gof.sources]
res=[None]*32
for _ in range(2): met.compute(srcs,res,srcs)
torch.cuda.synchronize(); t0=time.time()
for _ in range(3): met.compute(srcs,res,srcs)
torch.cuda.synchronize(); print("wall ms per gof", (time.time()-t0)/3*1e3)
codec.enableTiming(True); met.compute(srcs,res,srcs); t=codec.timings(); codec.enableTiming(False)
tot=sum(v[0] for v in t.values()); print("kernel total ms", tot)
for k,v in sorted(t.items(), key=lambda kv:-kv[1][0]): print(k, round(v[0],3), v[1])
# the same with the sources cached in HBM (what bench.py's transcode loop does)
dsrcs=[dict(positions=torch.from_numpy(s["positions"]).cuda(), colors=torch.from_numpy(s["colors"]).cuda(), normals=torch.from_numpy(s["normals"]).cuda()) for s in gof.sources]
for _ in range(2): met.compute(dsrcs,res,dsrcs)
torch.cuda.synchronize(); t0=time.time()
for _ in range(3): met.compute(dsrcs,res,dsrcs)
torch.cuda.synchronize(); print("resident sources: wall ms per gof", (time.time()-t0)/3*1e3)
codec.enableTiming(True); met.compute(dsrcs,res,dsrcs); t=codec.timings(); codec.enableTiming(False)
tot=sum(v[0] for v in t.values()); print("kernel total ms", tot)
for k,v in sorted(t.items(), key=lambda kv:-kv[1][0]): print(k, round(v[0],3), v[1])
