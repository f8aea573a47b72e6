"""short Rec-1 decode (8 vox10 frames) for ncu captures of the kd-forest kernels"""
import sys
sys.path.insert(0, '/root/repo')
import torch
import rabbit_transcoding_b200 as rb
kw = dict(bitdepth=10, width=1280, scale=0.68, height_blocks=80, seed=0x0AB817, transfer_filter=1, max_depth=249)
gof = rb.synthetic.generate_gof_parallel(8, workers=8, **kw)
codec = rb.codec.PCCCodecB200(device=0)
codec.uploadGof(gof)
for _ in range(2):
    codec.decodeGof()
torch.cuda.synchronize()
print("ok", sum(c.total for c in codec.frameCounts()))
