import os, sys, cProfile, pstats
sys.argv=['x']; os.environ['AMP']='1'; os.environ['NOPROF']='1'
ROOT='/root/repo'
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch, time
import torch.nn.functional as F
from helpers import edvr_state_shapes
from realvsr_b200.archs import EDVR_arch as E
from synth import synth_input, synth_state_dict
kw = dict(nf=64, nc=3, nframes=5, groups=8, front_RBs=5, back_RBs=10, w_TSA=True)
net = E.EDVR(**kw)
net.load_state_dict(synth_state_dict(edvr_state_shapes("EDVR", **kw), 7), strict=True)
net = net.to("cuda:0").train()
x = synth_input((16, 5, 3, 64, 64), 9).to("cuda:0")
gt = synth_input((16, 3, 256, 256), 10).to("cuda:0")
def step():
    net.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = F.l1_loss(net(x).float(), gt)
    loss.backward()
for _ in range(5): step()
torch.cuda.synchronize()
# host time alone: time to ISSUE a step with the GPU idle at start
ts=[]
for _ in range(5):
    torch.cuda.synchronize(); t=time.perf_counter(); step(); ts.append(time.perf_counter()-t)
print('host issue time per step (ms):', ['%.2f'%(1e3*t) for t in ts])
# forward only
ts=[]
for _ in range(5):
    torch.cuda.synchronize(); t=time.perf_counter()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = F.l1_loss(net(x).float(), gt)
    ts.append(time.perf_counter()-t); loss.backward()
print('forward issue time (ms):', ['%.2f'%(1e3*t) for t in ts])
pr=cProfile.Profile(); pr.enable()
for _ in range(5): step()
pr.disable(); torch.cuda.synchronize()
st=pstats.Stats(pr); st.sort_stats('tottime').print_stats(28)
