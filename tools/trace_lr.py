import sys; sys.path.insert(0,'.')
import torch, numpy as np
from se3_transformer_pytorch_b200 import ops
dev='cuda'; E=65536; Co=512
P=int(sys.argv[1]) if len(sys.argv)>1 else 1
Kp=int(sys.argv[2]) if len(sys.argv)>2 else 16
Ci=512; F=1
T=torch.randn(ops.t_numel(E//128,Ci,F,P),device=dev); U=torch.randn(E,64,device=dev)
img=ops.pack_lowrank(torch.randn(Co*Ci*F,Kp,device=dev),Co,Ci,F,Kp)
out=torch.empty(E,Co,P,device=dev)
trace=torch.zeros(5*64*8,dtype=torch.int64,device=dev)
lib=ops.lib()
for _ in range(2): ops.pairwise_lr(U,img,T,E,Co,Ci,F,P,Kp,out,False)
rc=lib.se3_pairwise_lr_trace(U.data_ptr(),img.data_ptr(),T.data_ptr(),E,Co,Ci,F,P,Kp,0,out.data_ptr(),trace.data_ptr(),torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize(); assert rc==0
t=trace.cpu().numpy().reshape(5,64,8)
t0=t[t>0].min()
rel=np.where(t>0,t-t0,-1)
names=['MMA','EPI4','EPI19','Wprod','Tprod']
for s in range(20,30):
    print('step',s, ' | '.join(f"{names[r]}:"+','.join(str(int(x)) for x in rel[r,s] if x>=0) for r in range(5)))
# per-step period
for r in range(5):
    d=np.diff(rel[r,10:60,0]); print(names[r],'period mean',d.mean(),'min',d.min(),'max',d.max())
