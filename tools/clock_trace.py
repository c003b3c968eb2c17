"""Run cfg2 forwards back to back for a few seconds and print the nvidia-smi clock / power trace next to the
step time of each 50-step chunk (is the step power-capped?)."""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
    sys.path.insert(0, p)
import torch
from helpers import edvr_state_shapes
from realvsr_b200.archs import EDVR_arch as E
from synth import synth_input, synth_state_dict
CFG = dict(nf=64, nc=3, nframes=5, groups=8, front_RBs=5, back_RBs=10, w_TSA=True)
net = E.EDVR(**CFG).eval()
net.load_state_dict(synth_state_dict(edvr_state_shapes("EDVR", **CFG), 7), strict=True)
net = net.to("cuda:0").half(); net.exec_path = "engine"
B = int(os.environ.get("B", "4"))
xs = [synth_input((B, 5, 3, 180, 320), 8 + i).to("cuda:0").half() for i in range(4)]
f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
with torch.no_grad():
    for i in range(3): net(xs[i % 4])
    torch.cuda.synchronize()
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,temperature.gpu,clocks_event_reasons.sw_power_cap",
                          "--format=csv,noheader,nounits", "-lms", "10", "-i", "0"], stdout=f)
    time.sleep(0.3)
    for chunk in range(int(os.environ.get("CHUNKS", "8"))):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(50): net(xs[i % 4])
        e1.record(); torch.cuda.synchronize()
        print("chunk %d: %.3f ms/step" % (chunk, e0.elapsed_time(e1) / 50))
    time.sleep(0.2)
    p.terminate(); p.wait()
f.flush(); f.seek(0)
lines = [l.strip() for l in f if l.strip()]
vals = [[float(t) for t in l.split(",")[:3]] for l in lines if l.split(",")[0].strip().isdigit()]
tail = vals[len(vals) // 2:-30] or vals
print("dbg=%s B=%d steady state: sm %.0f MHz, power %.0f W, temp %.0f C (%d samples)" % (
    os.environ.get("RVSR_TC_DEBUG", "0"), B, sum(v[0] for v in tail) / len(tail), sum(v[1] for v in tail) / len(tail),
    sum(v[2] for v in tail) / len(tail), len(tail)))
if os.environ.get("TRACE"):
    for l in lines[::4]:
        print(l)
