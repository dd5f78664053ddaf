"""Build the developer tuning harness tools/tune_tma (not part of the product library).

    python tools/build_tune.py

Links tools/tune_tma.cu against the objects levelsetpy_b200/build/*.o (run `python -m levelsetpy_b200.build` first).
"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from levelsetpy_b200 import build as _b  # noqa: E402


def main():
    _b.build()
    objs = [os.path.join(_b.OBJ, s.replace(".cu", ".o")) for s in _b.SOURCES]
    out = os.path.join(ROOT, "tools", "tune_tma")
    cmd = [_b.nvcc()] + _b.ARCH + ["-O3", "-std=c++20", "-lineinfo", "--expt-relaxed-constexpr", "-Xptxas", "-v",
                                   os.path.join(ROOT, "tools", "tune_tma.cu")] + objs + ["-o", out, "-cudart", "static", "-lcuda"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    with open(os.path.join(ROOT, "tools", "tune_tma.log"), "w") as fh:
        fh.write(r.stdout + r.stderr)
    if r.returncode != 0:
        print(r.stderr[-6000:])
        raise SystemExit(1)
    print(out)


if __name__ == "__main__":
    main()
