"""Rec-1 decode of a 32-frame vox10 GOF without the re-transfer (reconstruction + smoothing kernels only): for ncu"""
import sys
sys.path.insert(0, '/root/repo')
import torch
import rabbit_transcoding_b200 as rb
kw = dict(bitdepth=10, width=1280, scale=0.68, height_blocks=80, seed=0x0AB817, transfer_filter=0, max_depth=249)
gof = rb.synthetic.generate_gof_parallel(32, workers=16, **kw)
codec = rb.codec.PCCCodecB200(device=0)
codec.uploadGof(gof)
for _ in range(3):
    codec.decodeGof()
torch.cuda.synchronize()
