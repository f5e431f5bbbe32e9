import sys, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/mujoco-template_b200'); sys.path.insert(0,'/root/repo/tests')
import torch, numpy as np
from conftest import load_model, random_states
from mujoco_template import _capi, _mj as mj
model = load_model('cartpole'); n=65536
qpos,qvel,ctrl = random_states(model,'cartpole',n,seed=1)
data = mj.BatchData(model,n)
hq = torch.as_tensor(qpos.T.copy()).pin_memory(); hv = torch.as_tensor(qvel.T.copy()).pin_memory()
hu = torch.zeros((1,n),dtype=torch.float64).pin_memory(); hw = torch.zeros((2,n),dtype=torch.float64).pin_memory()
hA = torch.zeros((4,4,n),dtype=torch.float64).pin_memory(); hB = torch.zeros((4,1,n),dtype=torch.float64).pin_memory()
st = _capi.State(hq.data_ptr(), hv.data_ptr(), hu.data_ptr(), hw.data_ptr(), None)
b = data.backend.batch
def run(lin, A, B, reps=50):
    for _ in range(3): b.step_host(st,1,lin,1e-6,A,B,0)
    t0=time.perf_counter()
    for _ in range(reps): b.step_host(st,1,lin,1e-6,A,B,0)
    return (time.perf_counter()-t0)/reps*1e3
print('step_host lin+AB D2H ms', run(True, hA.data_ptr(), hB.data_ptr()))
print('step_host lin, no AB copy ms', run(True, None, None))
print('step_host no lin ms', run(False, None, None))
qn, vn, un = hq.numpy(), hv.numpy(), hu.numpy(); tmp=np.empty(n); K=np.random.rand(1,4)
def tick():
    rows=(qn[0],qn[1],vn[0],vn[1]); np.multiply(rows[0],-K[0,0],out=un[0])
    for k in range(1,4): np.multiply(rows[k],-K[0,k],out=tmp); np.add(un[0],tmp,out=un[0])
    np.clip(un[0],-200,200,out=un[0])
t0=time.perf_counter()
for _ in range(100): tick()
print('tick ms',(time.perf_counter()-t0)/100*1e3)
for mb in (1, 4, 13, 64):
    x = torch.empty(mb * 1024 * 1024, dtype=torch.uint8, device='cuda'); h = torch.empty(mb * 1024 * 1024, dtype=torch.uint8).pin_memory()
    for _ in range(3): h.copy_(x, non_blocking=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): h.copy_(x, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
    for _ in range(3): x.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(20): x.copy_(h, non_blocking=True)
    torch.cuda.synchronize(); dt2 = (time.perf_counter() - t0) / 20
    print(f'{mb} MB: D2H {dt*1e3:.3f} ms {mb/1024/dt:.1f} GB/s   H2D {dt2*1e3:.3f} ms {mb/1024/dt2:.1f} GB/s')
