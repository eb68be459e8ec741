"""Per-launch CUDA-event times of one forward (depth-1 slice of cfg2), in launch order.  Run on the GPU box."""
import sys; sys.path.insert(0, '.')
import torch
from se3_transformer_pytorch_b200 import SE3Transformer, ops
torch.manual_seed(0)
with torch.device('cuda'):
    m = SE3Transformer(dim=512, heads=8, depth=1, dim_head=64, num_degrees=4, num_neighbors=16, valid_radius=10).eval()
m.pack_weights(free_master=True, max_distance=16.0)
feats = torch.randn(4, 1024, 512, device='cuda'); coors = torch.randn(4, 1024, 3, device='cuda'); mask = torch.ones(4, 1024, dtype=torch.bool, device='cuda')
with torch.no_grad():
    for _ in range(2): m(feats, coors, mask)
    torch.cuda.synchronize()
    ops.PROFILE = []
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(); m(feats, coors, mask); t1.record(); torch.cuda.synchronize()
prof, ops.PROFILE = ops.PROFILE, None
print('forward ms', t0.elapsed_time(t1))
prev_end = t0
rows = []
for name, s, e, fl, nb, tag, ex in prof:
    rows.append((name, tag, prev_end.elapsed_time(s), s.elapsed_time(e)))
    prev_end = e
for r in rows[:90]:
    print(f'{r[0]:12s} {r[1]:28s} gap {r[2]:7.3f}  kernel {r[3]:7.3f}')
print('sum gaps', sum(r[2] for r in rows), 'sum kernels', sum(r[3] for r in rows))
