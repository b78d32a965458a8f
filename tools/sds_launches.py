"""two eager (no CUDA graph) fp32 SDS steps for an ncu launch list:
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/sds_launches.csv python tools/sds_launches.py [mode]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else 'fp32'
dev = torch.device('cuda:0')
z, emb = bench.make_guidance(dev, mode)
z.graph = False
pol, az, rad = torch.tensor([10.0]), torch.tensor([30.0]), torch.tensor([0.1])
for it in range(3):
    pred = torch.rand(1, 3, 72, 72, device=dev, requires_grad=True)
    z.train_step(emb, pred, pol, az, rad, guidance_scale=5, grad_scale=0.01)[0].backward()
    torch.cuda.synchronize()
print('done')
