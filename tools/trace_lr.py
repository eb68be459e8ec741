import sys; sys.path.insert(0, '.')
import torch, numpy as np
from se3_transformer_pytorch_b200 import ops
dev = 'cuda'; E = 65536; Co = 512
P = int(sys.argv[1]) if len(sys.argv) > 1 else 1
Kp = int(sys.argv[2]) if len(sys.argv) > 2 else 16
Ci = 512; F = 1
T = torch.randn(ops.t_numel(E // 128, Ci, F, P), device=dev); U = torch.randn(E, 64, device=dev)
img = ops.pack_lowrank(torch.randn(Co * Ci * F, Kp, device=dev), Co, Ci, F, Kp)
out = torch.empty(E, Co, P, device=dev)
trace = torch.zeros(5 * 64 * 8, dtype=torch.int64, device=dev)
lib = ops.lib()
for _ in range(2): ops.pairwise_lr(U, img, T, E, Co, Ci, F, P, Kp, out, False)
rc = lib.se3_pairwise_lr_trace(U.data_ptr(), img.data_ptr(), T.data_ptr(), E, Co, Ci, F, P, Kp, 0, out.data_ptr(), trace.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize(); assert rc == 0
t = trace.cpu().numpy().reshape(5, 64, 8).astype(np.int64)
names = ['MMA', 'EPI4', 'EPI19', 'Wprod', 'Tprod']
lo, hi = 16, 60
for r in range(5):
    per = np.diff(t[r, lo:hi, 0]).mean()
    evs = [e for e in range(8) if t[r, lo, e] > 0]
    segs = []
    for a, b in zip(evs[:-1], evs[1:]):
        segs.append(f'{a}->{b}:{(t[r, lo:hi, b] - t[r, lo:hi, a]).mean():.0f}')
    print(f'P{P} K{Kp} {names[r]:6s} period {per:.0f} | ' + ' '.join(segs))
# cross-role: epilogue warp 4 step start minus MMA issue end for the same step
print('EPI4 start(s) - MMA issued(s):', (t[1, lo:hi, 0] - t[0, lo:hi, 4]).mean(), ' MMA wait-begin(s+3) - EPI4 release(s):', (t[0, lo + 3:hi, 0] - t[1, lo:hi - 3, 2]).mean())
