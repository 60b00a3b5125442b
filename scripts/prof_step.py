"""One warm step of the bench workload between cudaProfilerStart/Stop, for
  ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum ...
(per-launch DRAM traffic of every kernel of the step -> profiles/<tag>_step_dram.json via make_profiles.py)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from genpercept_b200 import weights as W  # noqa: E402
from genpercept_b200.pipeline import GenPerceptPipeline  # noqa: E402

B, R = int(os.environ.get("B", 8)), int(os.environ.get("R", 768))
state = W.synth_state(1234, with_dpt=False)
pipe = GenPerceptPipeline(unet=state["unet"], vae=state["vae"], text_embed=bench.text_embed(), torch_dtype=torch.float16)
x = torch.randint(0, 256, (B, 3, R, R), dtype=torch.uint8, device="cuda")
out = torch.empty((B, 1, R, R), dtype=torch.float32, device="cuda")
pipe._ensure_ready()
for _ in range(3):
    pipe._engine.infer(x, out_channels=1, out=out)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
pipe._engine.infer(x, out_channels=1, out=out)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("step done")
