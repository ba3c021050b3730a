"""Developer check (gpurun): per-launch timeline of one warmed-up evaluation (MPIDB200_TRACE) and the cost of the
stage-timer events.  Usage: MPIDB200_TRACE=gpurun_out/trace.csv [MPIDB200_SOLVER=cg] python tools/trace_run.py [96k|1m|996]"""
import os
import sys
import time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mpidopenmmplugin_b200.workloads import water_box, make_kernel

tiles = {"996": (1, 1, 1), "96k": (4, 4, 2), "1m": (7, 7, 7)}[sys.argv[1] if len(sys.argv) > 1 else "96k"]
s = water_box(tiles, polarization=int(os.environ.get("MPIDB200_POL", "0")), epsilon=1e-5)      # 0 Mutual, 1 Direct, 2 Extrapolated
k = make_kernel(s, solver=os.environ.get("MPIDB200_SOLVER", "diis"))
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
pos = torch.tensor(s.pos, dtype=torch.float64, device="cuda")
f = torch.zeros((s.n, 3), dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
k.setStream(st.cuda_stream)
for _ in range(6):          # the 4th evaluation is traced
    k.execute_device(pos.data_ptr(), True, True, f.data_ptr())
for prof in (False, True, False, True):
    k.setProfiling(prof)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50):
        k.execute_device(pos.data_ptr(), True, True, f.data_ptr())
    torch.cuda.synchronize()
    print("profiling", prof, "ms/eval %.4f" % ((time.perf_counter() - t0)*1e3/50))
print("iterations", k.getStats()["iterations"], "eps", k.getStats()["epsilon"])
