#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_imgproc.py -q --no-header -p no:cacheprovider -s > gpurun_out/pytest_imgproc.log 2>&1
echo "imgproc tests exit $?"; grep -E "passed|failed|__call__" gpurun_out/pytest_imgproc.log | tail -n 4; grep -E "^FAILED|^E  " gpurun_out/pytest_imgproc.log | head -n 20
bash scripts/gpu_quick.sh
python - <<'PY'
# pre/post timings at a realistic size: 1080p photo -> 768-max-edge -> back, colour map
import time, torch, numpy as np
from genpercept_b200 import engine as E
from oracle.imgproc import spectral_lut_u8
x = torch.randint(0, 256, (1, 3, 1080, 1920), dtype=torch.uint8).pin_memory()
p = torch.rand((1, 1, 432, 768), device="cuda")
lut = spectral_lut_u8()
def t(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("resize 1080p u8 host -> 432x768 cuda: %.3f ms" % t(lambda: E.resize_aa(x, 432, 768, device="cuda")))
print("resize back f32 432x768 -> 1080p cuda: %.3f ms" % t(lambda: E.resize_aa(p, 1080, 1920)))
big = torch.rand((1, 1080, 1920), device="cuda")
print("colorize 1080p -> host u8 HWC: %.3f ms" % t(lambda: E.colorize(big, lut)))
print("quantize16 1080p -> host: %.3f ms" % t(lambda: E.quantize(big, 16)))
from torchvision.transforms.functional import resize
torch.set_num_threads(16)
t0 = time.perf_counter(); [resize(x, [432, 768], antialias=True) for _ in range(5)]; print("torchvision CPU resize 1080p->432x768: %.2f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
PY
