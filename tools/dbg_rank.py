import sys; sys.path.insert(0,'.')
import torch
from se3_transformer_pytorch_b200 import SE3Transformer, ops
torch.manual_seed(0)
dev='cuda'
with torch.device(dev):
    m = SE3Transformer(dim=128, heads=8, dim_head=16, depth=1, num_degrees=4, num_neighbors=16).eval()
conv = m.net.blocks[0][0].attn.to_v
coors = torch.randn(4,1024,3,device=dev)
idx, nm, rp, rd = ops.knn(coors, 16, 1e5)
feat = rd.reshape(-1,1).contiguous()
pk = conv.packed()
G = ops.radial_trunk(feat, pk['trunk'], len(conv.pairs))
# fp64 recomputation of the trunk
def trunk64(pc, x):
    n = pc.rp.net
    x = x.double()
    h = torch.nn.functional.linear(x, n['0'].weight.double(), n['0'].bias.double())
    h = torch.nn.functional.gelu(torch.nn.functional.layer_norm(h, (128,), n['1'].weight.double(), n['1'].bias.double()))
    h = torch.nn.functional.linear(h, n['3'].weight.double(), n['3'].bias.double())
    return torch.nn.functional.gelu(torch.nn.functional.layer_norm(h, (128,), n['4'].weight.double(), n['4'].bias.double()))
for pi,(di,do) in enumerate(conv.pairs[:6]):
    pc = conv.kernel_unary[f'({di},{do})']
    G64 = trunk64(pc, feat)
    gmax = float(G64.abs().max())
    noise = float((G[pi].double()-G64).abs().max())/gmax
    line=[f'pair ({di},{do}) gmax {gmax:.2f} fp32-noise {noise:.1e}']
    for src,name in ((G[pi].double(),'G32'),(G64,'G64')):
        C = src.t()@src
        ev,evec = torch.linalg.eigh(C)
        for r in (15,23,31,47):
            V = evec[:,128-r:]
            res = float((src-(src@V)@V.t()).abs().max())/gmax
            line.append(f'{name} r{r}:{res:.1e}')
    print(' '.join(line))
print('dist range', float(rd.min()), float(rd.max()))
