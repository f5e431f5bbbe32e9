"""FP32 mode: trajectory divergence from the FP64 path over 1000 steps (printed, then pinned in tests)."""
import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/mujoco-template_b200'); sys.path.insert(0,'/root/repo/tests')
import torch, numpy as np
from conftest import load_model, random_states
from mujoco_template import _mj as mj
for name in ('pendulum','cartpole','drone','humanoid'):
    model = load_model(name); n = 256
    qpos,qvel,ctrl = random_states(model,name,n,seed=21)
    if name=='pendulum': ctrl[:]=0; qvel[:]=0
    if name=='cartpole': ctrl[:]=0
    if name=='drone': ctrl[:]=3.2495625
    outs={}
    for prec in (64,32):
        d = mj.BatchData(model, n, precision=prec)
        dt = d.qpos.dtype
        d.qpos.copy_(torch.as_tensor(qpos.T.copy(),device='cuda').to(dt)); d.qvel.copy_(torch.as_tensor(qvel.T.copy(),device='cuda').to(dt)); d.ctrl.copy_(torch.as_tensor(ctrl.T.copy(),device='cuda').to(dt))
        traj=[]
        for k in range(10):
            mj.mj_step(model,d,100); traj.append(d.qpos.double().cpu().numpy().copy())
        outs[prec]=(np.array(traj), int(d.flags.max()), d.backend.batch.kernel_variant)
    diff=np.abs(outs[64][0]-outs[32][0])
    print(name, 'variants', outs[64][2], outs[32][2], 'flags', outs[64][1], outs[32][1], 'max |dq| at 100..1000 steps:', ['%.2e'%diff[k].max() for k in (0,4,9)], 'median', '%.2e'%np.median(diff[9].max(axis=0)))
