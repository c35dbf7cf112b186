"""Developer tool: host timeline milestones of replayed forwards (SCN_TIMELINE=1 is set here)."""
import os
import re
import subprocess
import sys

if os.environ.get("SCN_TIMELINE") != "1":
    env = dict(os.environ, SCN_TIMELINE="1")
    p = subprocess.run([sys.executable, __file__] + sys.argv[1:], env=env, capture_output=True, text=True)
    fw, cur = [], []
    for line in p.stderr.splitlines():
        if line.startswith("[tl]"):
            m = re.match(r"\[tl\]\s+([\d.]+) us (\S+)\s+kind (\d+)\s+(-?\d+) (-?\d+)", line)
            cur.append((float(m.group(1)), m.group(2), int(m.group(3)), int(m.group(4)), int(m.group(5))))
        elif line.startswith("forward"):
            fw.append((line, cur))
            cur = []
        else:
            print(line)
    if "--raw" in sys.argv:  # every milestone of the last forward: time, thread role, kind, two arguments
        for e in fw[-1][1]:
            print("%9.1f us  %-7s kind %d  %d %d" % e)
    for line, evs in fw[-6:]:
        def first(pred):
            return next((e[0] for e in evs if pred(e)), float("nan"))
        last_worker = max([e[0] for e in evs if e[1] == "worker"], default=float("nan"))
        last_main = max([e[0] for e in evs if e[1] == "main"], default=float("nan"))
        print(line, "| input built %.0f | subm L0 ready %.0f | conv L0->L1 %.0f | conv L1->L2 %.0f | deepest conv %.0f | worker done %.0f | main done %.0f us" % (
            first(lambda e: e[1] == "input" and e[2] == 1), first(lambda e: e[1] == "worker" and e[2] == 1), first(lambda e: e[1] == "worker" and e[2] == 2),
            first(lambda e: e[1] == "worker" and e[2] == 2 and e[3] == 1024), first(lambda e: e[1] == "worker" and e[2] == 2 and e[3] == 16), last_worker, last_main))
    sys.exit(0)

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import fpn_util  # noqa: E402
import detection_3d_b200.sparseconvnet as scn  # noqa: E402
from detection_3d_b200 import synthetic  # noqa: E402

scn.set_math_mode(next((a for a in sys.argv[1:] if not a.startswith("--")), "bf16"))
net = scn.FPN_Net(**scn.sw4c_fpn432_config())
net.load_state_dict(fpn_util.deterministic_state(net, seed=1))
net = net.cuda().eval()
coords = torch.from_numpy(synthetic.building_coords()).cuda()
feats = torch.from_numpy(fpn_util.features_for(coords.cpu().numpy())).cuda()
with torch.no_grad():
    for i in range(12):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = net([coords, feats])
        e1.record()
        torch.cuda.synchronize()
        del out
        print(f"forward {i}: {e0.elapsed_time(e1):.2f} ms", file=sys.stderr, flush=True)
