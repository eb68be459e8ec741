import sys; sys.path.insert(0, '.')
import os, threading, time
import torch
import pynvml
from se3_transformer_pytorch_b200 import ops
dev = 'cuda'
E = 65536; Co = 512; Ci = 512
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)

class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True); self.stop = False; self.clk = []; self.pw = []
    def run(self):
        while not self.stop:
            self.clk.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)); self.pw.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1e3)
            time.sleep(0.004)

def run(P, F, Kp, secs=0.5):
    T = torch.randn(ops.t_numel(E // 128, Ci, F, P), device=dev)
    U = torch.randn(E, 64, device=dev)
    Fp = torch.randn(Co * Ci * F, Kp, device=dev)
    img = ops.pack_lowrank(Fp, Co, Ci, F, Kp)
    out = torch.empty(E, Co, P, device=dev)
    for _ in range(2): ops.pairwise_lr(U, img, T, E, Co, Ci, F, P, Kp, out, False)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); ops.pairwise_lr(U, img, T, E, Co, Ci, F, P, Kp, out, False); e1.record(); torch.cuda.synchronize()
    n = max(3, int(secs * 1e3 / e0.elapsed_time(e1)))
    s = Sampler(); s.start()
    e0.record()
    for _ in range(n): ops.pairwise_lr(U, img, T, E, Co, Ci, F, P, Kp, out, False)
    e1.record(); torch.cuda.synchronize()
    s.stop = True; s.join()
    k = len(s.clk) // 3
    clk = sorted(s.clk[k:])[len(s.clk[k:]) // 2] if s.clk[k:] else 0
    pw = sorted(s.pw[k:])[len(s.pw[k:]) // 2] if s.pw[k:] else 0
    return e0.elapsed_time(e1) / n, clk, pw
res = []
cases = ((1, 1), (3, 3), (5, 5), (7, 7)) if not os.environ.get('MICRO_CASES') else [tuple(map(int, c.split('x'))) for c in os.environ['MICRO_CASES'].split(',')]
for P, F in cases:
    for Kp in (16, 32):
        ms, clk, pw = run(P, F, Kp)
        steps = (E // 128) * (Co // 32) * (Ci * F // 4) / 148
        res.append(f'P{P}F{F}K{Kp}={ms:.2f}@{clk}MHz/{pw:.0f}W/{ms * clk * 1e3 / steps:.0f}cyc')
print(' '.join(res))
