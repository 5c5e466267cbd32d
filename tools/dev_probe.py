import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from abb_dualarm_planning_b200 import Checker, sampling
n = 1_000_000
ck = Checker(1)
q = sampling.sample_gd(1, 0, n)
dq = torch.from_numpy(q).cuda().t().contiguous(); dqo = torch.empty((12, n), dtype=torch.float64, device="cuda"); dv = torch.empty(n, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
def step(): ck.gd_validate_dev(n, dq.data_ptr(), dqo.data_ptr(), dv.data_ptr(), st)
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): step()
e1.record(); torch.cuda.synchronize()
print("overlap on: %.3f ms/step" % (e0.elapsed_time(e1) / 5), flush=True)
ck.set_stage_timing(True); step(); torch.cuda.synchronize()
ck.L.ckc_debug_timeline(ck.ctx)
