"""Build libs3od_b200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m s3od_b200.build [--force]

nvcc cross-compiles without a GPU; the .so is git-ignored but travels with the tree to the GPU box.
"""
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB_DIR = os.path.join(HERE, "_lib")
LIB_PATH = os.path.join(LIB_DIR, "libs3od_b200.so")
SOURCES = ["engine.cu", "kernels_gemm_enc.cu", "kernels_gemm_head.cu", "kernels_misc.cu", "train.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "--expt-relaxed-constexpr"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _newest_source_mtime():
    m = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(root):
            m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


def build_lib(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIB_DIR, exist_ok=True)
    if not force and os.path.exists(LIB_PATH) and os.path.getmtime(LIB_PATH) >= _newest_source_mtime():
        print(f"s3od_b200.build: {LIB_PATH} is newer than every source under csrc/ and include/ - reused (force=True / --force recompiles)")
        return LIB_PATH
    print(f"s3od_b200.build: compiling {len(SOURCES)} translation units for sm_100a with nvcc ({'forced' if force else 'sources changed'})")
    nvcc = _nvcc()
    extra = os.environ.get("S3OD_NVCC_FLAGS", "").split()       # experiments only (e.g. -DS3OD_ATTN_POLY_EVERY=4)
    objs = [os.path.join(LIB_DIR, s.replace(".cu", ".o")) for s in SOURCES]

    def compile_one(args):
        src, obj = args
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose and (r.stdout or r.stderr):
            print(r.stdout, r.stderr)

    with cf.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        list(ex.map(compile_one, zip(SOURCES, objs)))
    cmd = [nvcc, "-shared", "-o", LIB_PATH] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose=True))
