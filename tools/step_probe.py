"""A few eager training steps of a bench workload and nothing else — the command to put under ncu for a launch list
(bench.py itself also generates label maps, captures graphs and times micro-benchmarks, all of which ncu would profile).

    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/step_probe.py [workload] [steps]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from favit_b200.engine import TrainStep

name = sys.argv[1] if len(sys.argv) > 1 else "vitb16_mhla_224"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
wl = bench.WORKLOADS[name]
dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
model = bench.build_model(wl, dev)
if wl["kind"] == "sppp":
    model.validate_slots = False
step = TrainStep(model, process_group=None, cuda_graph=False)
batch = bench.make_batch(wl, wl["B"], seed=1234, device=dev)
torch.cuda.synchronize()
print("PROBE_BEGIN", flush=True)
for _ in range(steps):
    loss = step(*batch)
torch.cuda.synchronize()
print("PROBE_END", float(loss))
