import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/mujoco-template_b200'); sys.path.insert(0,'/root/repo/tests')
import torch
from conftest import load_model, random_states
from mujoco_template import _mj as mj
model = load_model('humanoid'); n = 4096
d = mj.BatchData(model, n)
qpos,qvel,ctrl = random_states(model,'humanoid',n,seed=0)
d.qpos.copy_(torch.as_tensor(qpos.T.copy(),device='cuda')); d.qvel.copy_(torch.as_tensor(qvel.T.copy(),device='cuda')); d.ctrl.copy_(torch.as_tensor(ctrl.T.copy(),device='cuda'))
for _ in range(4):
    mj.mj_step(model,d)
torch.cuda.synchronize(); print('ok', d.backend.batch.kernel_variant)
