import sys, time, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/mujoco-template_b200'); sys.path.insert(0,'/root/repo/tests')
import torch, numpy as np
from conftest import load_model, random_states
from mujoco_template import _mj as mj
model = load_model('humanoid')
for n in (2048, 16384):
    d = mj.BatchData(model, n)
    print('variant', d.backend.batch.kernel_variant)
    qpos,qvel,ctrl = random_states(model,'humanoid',n,seed=0)
    d.qpos.copy_(torch.as_tensor(qpos.T.copy(),device='cuda')); d.qvel.copy_(torch.as_tensor(qvel.T.copy(),device='cuda')); d.ctrl.copy_(torch.as_tensor(ctrl.T.copy(),device='cuda'))
    mj.mj_step(model,d); torch.cuda.synchronize()
    for steps in (1, 10):
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record(); mj.mj_step(model,d,steps); e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1); print(n, 'steps',steps,'ms',ms,'env-steps/s %.3g'%(n*steps/ms*1e3), 'ncon mean', float(d.ncon.float().mean()), 'nefc', float(d.nefc.float().mean()), 'iter', float(d.solver_iter.float().mean()), 'flags', int(d.flags.max()))
