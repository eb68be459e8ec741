"""Print the parity error of every golden model case (tensor-core and SIMT paths) -- run on the GPU box."""
import os, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests'); sys.path.insert(0, 'tests/golden')
import numpy as np, torch
from helpers import MODEL_CASES, BIG_CASES, load_case, rel_err, case_inputs, case_outputs
from detfill import fill_state_dict
from se3_transformer_pytorch_b200 import SE3Transformer

def run(name, simt, lowrank=False):
    if simt: os.environ['SE3B200_FORCE_SIMT'] = '1'
    else: os.environ.pop('SE3B200_FORCE_SIMT', None)
    if lowrank:
        os.environ['SE3B200_LOWRANK_MIN_EDGES'] = '0'; os.environ.pop('SE3B200_NO_LOWRANK', None)
    else:
        os.environ['SE3B200_NO_LOWRANK'] = '1'
    z, cfg = load_case(name)
    m = SE3Transformer(**cfg['ctor']); fill_state_dict(m, seed=11); m = m.cuda().eval()
    feats, coors, mask, extra = case_inputs(z)
    t = lambda a: torch.from_numpy(a).cuda()
    feats = {k: t(v) for k, v in feats.items()} if isinstance(feats, dict) else t(feats)
    out = m(feats, t(coors), t(mask), **{k: t(v) for k, v in extra.items()}, **cfg['fwd'])
    ref = case_outputs(z)
    if isinstance(ref, dict):
        return max(rel_err(out[d].cpu().numpy(), r) for d, r in ref.items())
    return rel_err(out.cpu().numpy(), ref)

print('case,rel_err_tensor_core_path,rel_err_simt_path,rel_err_lowrank_path')
for name in MODEL_CASES + BIG_CASES:
    print(f'{name},{run(name, False):.3e},{run(name, True):.3e},{run(name, False, True):.3e}')
