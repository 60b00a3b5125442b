"""What a 256 x 128 output tile costs in the staged epilogue: the K-packed stem (K = 1 chunk: nothing but epilogue) and the
128->128 3x3 convolution (K = 18 chunks), each plain / with GroupNorm statistics / with a residual / with the direct
epilogue.  usage: python scripts/epilogue_probe.py"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = ("import sys, torch; sys.path.insert(0, %r); from genpercept_b200 import engine as E; "
        "us, fl = E.bench_conv(torch.float16, *%%s, iters=20); print('%%s: %%.0f us  %%.0f TFLOP/s' %% (%%r, us, fl / us / 1e6))" % ROOT)
SHAPES = {"stem 32->128 1x1 @768^2 x8": (8, 768, 768, 32, 128, 1, 0), "conv 128->128 3x3 @768^2 x8": (8, 768, 768, 128, 128, 3, 0),
          "conv 256->256 3x3 @384^2 x8": (8, 384, 384, 256, 256, 3, 0)}
VARIANTS = {"plain": {}, "stats": {"GP_BENCH_STATS": "1"}, "residual": {"GP_BENCH_RES": "1"},
            "stats+residual": {"GP_BENCH_STATS": "1", "GP_BENCH_RES": "1"}, "residual, no L2 prefetch": {"GP_BENCH_RES": "1", "GP_NO_RES_PREFETCH": "1"},
            "residual, per-thread loads": {"GP_BENCH_RES": "1", "GP_NO_RES_TMA": "1"}, "direct epilogue": {"GP_DIRECT_EPILOGUE": "1"}}
for sname, shape in SHAPES.items():
    for vname, env in VARIANTS.items():
        code = ("import sys, torch; sys.path.insert(0, %r); from genpercept_b200 import engine as E; us, fl = E.bench_conv(torch.float16, *%r, iters=20); "
                "print('%-30s %-28s %%7.0f us  %%6.0f TFLOP/s' %% (us, fl / us / 1e6))" % (ROOT, shape, sname, vname))
        p = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True)
        print(p.stdout.strip() or p.stderr.strip()[-300:])
