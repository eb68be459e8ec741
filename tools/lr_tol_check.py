"""Headline-width check of the low-rank plan tolerance: forward of a depth-2 slice of cfg2 (b=1) with the direct K=128
tensor-core kernel vs the low-rank path at several plan tolerances.  Run on the GPU box."""
import os, sys, collections
sys.path.insert(0, '.')
import torch
from se3_transformer_pytorch_b200 import SE3Transformer

def build():
    torch.manual_seed(0)
    with torch.device('cuda'):
        m = SE3Transformer(dim=512, heads=8, depth=2, dim_head=64, num_degrees=4, num_neighbors=16, valid_radius=10).eval()
    return m

torch.manual_seed(1)
feats = torch.randn(1, 1024, 512, device='cuda'); coors = torch.randn(1, 1024, 3, device='cuda') * 3; mask = torch.ones(1, 1024, dtype=torch.bool, device='cuda')
os.environ['SE3B200_NO_LOWRANK'] = '1'
with torch.no_grad():
    ref = build()(feats, coors, mask).double()
del os.environ['SE3B200_NO_LOWRANK']
os.environ['SE3B200_FORCE_SIMT'] = '1'; os.environ['SE3B200_NO_LOWRANK'] = '1'
with torch.no_grad():
    simt = build()(feats, coors, mask).double()
del os.environ['SE3B200_FORCE_SIMT'], os.environ['SE3B200_NO_LOWRANK']
print(f'direct tensor-core vs fp32 SIMT: {float((ref - simt).abs().max() / simt.abs().max()):.2e}')
os.environ['SE3B200_LOWRANK_MIN_EDGES'] = '0'
for tol in ('2e-7', '1e-6', '1.5e-6', '3e-6', '1e-5'):
    os.environ['SE3B200_LOWRANK_TOL'] = tol
    m = build()
    with torch.no_grad():
        out = m(feats, coors, mask).double()
    ks = collections.Counter()
    for mod in m.modules():
        pk = getattr(mod, '_packed', None)
        if pk and pk.get('lr'):
            for v in pk['lr']['pairs'].values():
                ks[v['Kp']] += 1
    print(f'tol {tol}: vs direct {float((out - ref).abs().max() / ref.abs().max()):.2e}  vs SIMT {float((out - simt).abs().max() / simt.abs().max()):.2e}  Kp histogram {dict(ks)}')
