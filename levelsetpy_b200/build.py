"""Build the CUDA shared library in-tree: levelsetpy_b200/_hjb200.so  (sm_100a only).

    python -m levelsetpy_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the snapshot.
"""
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
SO = os.path.join(HERE, "_hjb200.so")
SOURCES = ["hj_api.cu", "hj_gather.cu", "hj_tma.cu", "hj_halo.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++20", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def nvcc():
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found")
    return exe


def _deps_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(root):
            m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def _compile(src, verbose, extra=(), objdir=OBJ):
    obj = os.path.join(objdir, src.replace(".cu", ".o"))
    cmd = [nvcc()] + ARCH + FLAGS + list(extra) + ["-c", os.path.join(CSRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(objdir, src + ".log")
    with open(log, "w") as fh:
        fh.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stderr[-6000:]))
    if verbose:
        print(r.stderr)
    return obj


def build(force=False, verbose=False):
    if not force and os.path.exists(SO) and os.path.getmtime(SO) >= _deps_mtime():
        return SO
    os.makedirs(OBJ, exist_ok=True)
    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(lambda s: _compile(s, verbose), SOURCES))
    return _link(objs, SO)


def _link(objs, so):
    cmd = [nvcc()] + ARCH + ["-shared", "-o", so] + objs + ["-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stderr[-4000:])
    return so


def build_variant(tag, defines):
    """Developer tooling (tools/build_variants.py): the library with other tile shapes, as _hjb200_<tag>.so.  Only
    hj_tma.cu depends on the tile-shape macros; the other objects are the production ones."""
    build()
    vdir = os.path.join(OBJ, "variant_" + tag)
    os.makedirs(vdir, exist_ok=True)
    obj = _compile("hj_tma.cu", False, ["-D%s" % d for d in defines], vdir)
    objs = [obj if s == "hj_tma.cu" else os.path.join(OBJ, s.replace(".cu", ".o")) for s in SOURCES]
    return _link(objs, os.path.join(HERE, "_hjb200_%s.so" % tag))


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
