import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/mujoco-template_b200'); sys.path.insert(0,'/root/repo/tests')
import torch, numpy as np
n=65536
h = torch.empty((20, n), dtype=torch.float64).pin_memory(); d = torch.empty((20,n), dtype=torch.float64, device='cuda')
hs = torch.empty((7, n), dtype=torch.float64).pin_memory(); ds = torch.empty((7,n), dtype=torch.float64, device='cuda')
torch.cuda.synchronize()
def t(fn, reps=20):
    fn(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/reps*1e3
print('D2H 10.5MB ms', t(lambda: h.copy_(d, non_blocking=True)), 'GB/s', 20*n*8/1e9/(t(lambda: h.copy_(d, non_blocking=True))*1e-3))
print('H2D 3.7MB ms', t(lambda: ds.copy_(hs, non_blocking=True)))
K=np.random.rand(1,4); hq=np.random.rand(2,n); hv=np.random.rand(2,n); hu=np.zeros((1,n))
def tick():
    x=np.concatenate([hq,hv],axis=0); hu[:] = np.clip(-(K@x),-200,200)
t0=time.perf_counter()
for _ in range(50): tick()
print('host LQR tick ms', (time.perf_counter()-t0)/50*1e3)
s1=torch.cuda.Stream(); s2=torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2): ds.copy_(hs, non_blocking=True)
print('D2H+H2D concurrent ms', t(both))
