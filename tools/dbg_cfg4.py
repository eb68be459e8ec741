import os, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'tests/golden')
import numpy as np, torch
from helpers import load_case, rel_err, case_inputs, case_outputs
from detfill import fill_state_dict
from se3_transformer_pytorch_b200 import SE3Transformer

name = sys.argv[1] if len(sys.argv) > 1 else 'cfg4_b2'
z, cfg = load_case(name)
def run(simt):
    if simt: os.environ['SE3B200_FORCE_SIMT'] = '1'
    else: os.environ.pop('SE3B200_FORCE_SIMT', None)
    m = SE3Transformer(**cfg['ctor']); fill_state_dict(m, seed=11); m = m.cuda().eval()
    feats, coors, mask, extra = case_inputs(z)
    t = lambda a: torch.from_numpy(a).cuda()
    cap = {}
    hooks = [m.conv_in.register_forward_hook(lambda mod, a, o: cap.update(conv_in=o))]
    for L, blk in enumerate(m.net.blocks):
        hooks.append(blk[0].register_forward_hook(lambda mod, a, o, L=L: cap.update({f'attn{L}': o})))
        hooks.append(blk[1].register_forward_hook(lambda mod, a, o, L=L: cap.update({f'ff{L}': o})))
    out = m(t(feats), t(coors), t(mask), **{k: t(v) for k, v in extra.items()}, **cfg['fwd'])
    return out, cap
ref = case_outputs(z)
o_tc, c_tc = run(False)
o_si, c_si = run(True)
print('tc vs fixture', rel_err(o_tc.cpu().numpy(), ref), ' simt vs fixture', rel_err(o_si.cpu().numpy(), ref), ' tc vs simt', rel_err(o_tc.cpu().numpy(), o_si.cpu().numpy()))
for k in c_tc:
    for d in c_tc[k]:
        a, b = c_tc[k][d].cpu().numpy(), c_si[k][d].cpu().numpy()
        print(k, d, 'tc vs simt rel', rel_err(a, b), 'absmax', np.abs(b).max())
