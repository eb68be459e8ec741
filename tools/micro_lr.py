import sys; sys.path.insert(0,'.')
import torch, time
from se3_transformer_pytorch_b200 import ops
dev='cuda'
E=65536; Co=512
def run(P, Ci, F, Kp, n=5):
    di = (F-1)//2; do=(P-1)//2
    PH=(P+3)//4
    T = torch.randn(ops.t_numel(E//128, Ci, F, P), device=dev)
    U = torch.randn(E,64,device=dev)
    Fp = torch.randn(Co*Ci*F, Kp, device=dev)
    img = ops.pack_lowrank(Fp, Co, Ci, F, Kp)
    out = torch.empty(E,Co,P,device=dev)
    for _ in range(2): ops.pairwise_lr(U,img,T,E,Co,Ci,F,P,Kp,out,False)
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): ops.pairwise_lr(U,img,T,E,Co,Ci,F,P,Kp,out,False)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n
for P in (1,7):
    res=[]
    for Ci in (32,128,512,1024):
        ms=run(P,Ci,1,16)
        nifb=Ci//4
        res.append((nifb, ms))
        print(f'P={P} Ci={Ci} NIFB={nifb} K16: {ms:.3f} ms  per item-wave {ms/55.35*1e3:.1f} us')
    (n0,t0),(n1,t1)=res[1],res[3]
    b=(t1-t0)/55.35/(n1-n0)*1e6; a=t0/55.35*1e3-n0*b/1e3
    print(f'   => per step {b:.0f} ns, per item overhead {a:.1f} us')
