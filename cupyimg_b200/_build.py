"""Ahead-of-time build of libsepfilt_b200.so (sm_100a only, in-tree).

``python -m cupyimg_b200._build`` or ``cupyimg_b200._build.build()``; also called by
``__graft_entry__.build()``.  nvcc cross-compiles without a GPU.  The shared library
lands in ``cupyimg_b200/_lib/`` (git-ignored, travels to the GPU box with the snapshot).
"""
import hashlib
import os
import subprocess
import sys

_PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_PKG, "csrc")
LIB_DIR = os.path.join(_PKG, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libsepfilt_b200.so")
SOURCES = ["api.cu", "correlate_nd.cu", "exact.cu", "exact_tiled.cu", "exact_stream.cu", "minmax_stream.cu", "f32_1d.cu", "f32_stream.cu", "fused3d.cu", "fused_ws.cu", "fused_ws_p1.cu", "fused_ws_p2.cu", "fused_ws_w1.cu", "fused_ws_w2.cu", "fused_ws_g.cu", "consumers.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++20",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.sep not in cand or os.path.exists(cand)):
            return cand
    raise RuntimeError("nvcc not found")


HASH_PATH = os.path.join(LIB_DIR, "source.sha256")


def _source_hash():
    """Digest of every source the library is built from (mtimes do not survive a snapshot)."""
    h = hashlib.sha256()
    deps = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC))
    deps.append(os.path.join(_PKG, "..", "include", "sepfilt.h"))
    for d in deps:
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _stale():
    if not os.path.exists(LIB_PATH) or not os.path.exists(HASH_PATH):
        return True
    with open(HASH_PATH) as f:
        return f.read().strip() != _source_hash()


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ for sm_100a and link libsepfilt_b200.so."""
    if not force and not _stale():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    obj_dir = os.path.join(LIB_DIR, "obj")
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = _nvcc()
    procs = []
    objs = []
    for src in SOURCES:
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = False
    for src, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0 or verbose:
            sys.stderr.write("[%s]\n%s\n" % (src, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    tmp = LIB_PATH + ".tmp"
    subprocess.check_call([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", tmp, *objs,
                           "-Xlinker", "--no-undefined"])
    os.replace(tmp, LIB_PATH)
    with open(HASH_PATH, "w") as f:
        f.write(_source_hash())
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
