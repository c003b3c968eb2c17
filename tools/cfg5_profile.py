"""Where the cfg5 training step spends its time (torch profiler, CUDA kernels grouped by name)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch
import torch.nn.functional as F
from helpers import edvr_state_shapes
from realvsr_b200.archs import EDVR_arch as E
from synth import synth_input, synth_state_dict
kw = dict(nf=64, nc=3, nframes=5, groups=8, front_RBs=5, back_RBs=10, w_TSA=True)
net = E.EDVR(**kw)
net.load_state_dict(synth_state_dict(edvr_state_shapes("EDVR", **kw), 7), strict=True)
net = net.to("cuda:0").train()
if os.environ.get("CL", "0") == "1":   # experiment: channels_last weights / activations for the cuDNN convolutions
    net = net.to(memory_format=torch.channels_last)
x = synth_input((16, 5, 3, 64, 64), 9).to("cuda:0")
gt = synth_input((16, 3, 256, 256), 10).to("cuda:0")
amp = os.environ.get("AMP", "0") == "1"
def step():
    net.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
        loss = F.l1_loss(net(x).float(), gt)
    loss.backward()
if os.environ.get('GRAPH', '0') == '1':
    from realvsr_b200.train_c8 import GraphedStep
    gs = GraphedStep(net, F.l1_loss, x, gt, amp_dtype=torch.bfloat16 if amp else None)
    step = lambda: gs(x, gt)
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): step()
e1.record(); torch.cuda.synchronize()
print('cfg5 step: %.2f ms (AMP=%s, RVSR_TRAIN_C8=%s, CL=%s, GRAPH=%s)' % (e0.elapsed_time(e1) / 5, os.environ.get('AMP', '0'), os.environ.get('RVSR_TRAIN_C8', '1'), os.environ.get('CL', '0'), os.environ.get('GRAPH', '0')))
if os.environ.get('NOPROF', '0') == '1': sys.exit(0)
from torch.profiler import profile, ProfilerActivity
if os.environ.get('ATEN', '0') == '1':   # which torch operators (with shapes) are still on the step: eager only
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=True) as prof:
        step()
        torch.cuda.synchronize()
    rows = [e for e in prof.key_averages(group_by_input_shape=True) if e.key.startswith('aten::') and e.self_device_time_total > 0]
    rows.sort(key=lambda e: -e.self_device_time_total)
    for e in rows[:40]:
        print('%-34s n=%3d  %8.1f us  %s' % (e.key, e.count, e.self_device_time_total, str(e.input_shapes)[:150]))
    sys.exit(0)
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3): step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=int(os.environ.get('ROWS', '18')), max_name_column_width=70))
