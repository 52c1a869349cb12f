import sys; sys.path.insert(0,'/root/repo')
import torch
from d3feat.pytorch_b200 import _lib
lib=_lib.load()
n=24316320
p=torch.randn(n,device='cuda'); g=torch.randn(n,device='cuda'); m=torch.zeros(n,device='cuda')
lr=torch.full((1,),0.01,device='cuda'); flag=torch.zeros(1,dtype=torch.int32,device='cuda')
s=torch.cuda.current_stream().cuda_stream
for z in (0,1):
  for chk in (0,1):
    for _ in range(3): lib.d3f_sgd_step(p.data_ptr(),g.data_ptr(),m.data_ptr(),n,lr.data_ptr(),0.98,1e-6,flag.data_ptr(),chk,z,s)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): lib.d3f_sgd_step(p.data_ptr(),g.data_ptr(),m.data_ptr(),n,lr.data_ptr(),0.98,1e-6,flag.data_ptr(),chk,z,s)
    e1.record(); torch.cuda.synchronize()
    print('zero',z,'check',chk,'us/step',e0.elapsed_time(e1)*100)
