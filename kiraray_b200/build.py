#!/usr/bin/env python3
"""In-tree build of the product libraries (sm_100a only).

  kiraray_b200/lib/libkrr_wfpt.so  CUDA kernels + C ABI (include/krr_wfpt.h)          [nvcc]
  kiraray_b200/lib/libkrr_host.so  C++17 host layer (RenderPass surface, importers)  [g++]
  kiraray_b200/lib/krr_render      headless CLI driver (host/krr_render.cpp; same JSON configs as the reference)  [g++]

Input data: kiraray_b200/data/spectral_srgb.bin (colour-space tables, see data/MANIFEST.json) is NOT built here; check_data()
verifies it against the manifest's checksum and fails loudly when it is absent or different.

nvcc cross-compiles without a GPU.  The .so files are git-ignored but travel to the GPU box.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "lib")
OBJ = os.path.join(LIB, "obj")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CXX = os.environ.get("CXX", "g++")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
# -prec-div/-prec-sqrt=false: `/` and sqrtf() compile to the 2-ulp / 1-ulp fast sequences (the reference
# itself builds its device code with --use_fast_math, CMakeLists.txt:113).  Everything that is compared
# BIT-FOR-BIT with the oracle (sampler, wavelengths, camera rays, ray/triangle test) is written with the
# explicitly rounded intrinsics of krr_math.cuh (__fdiv_rn, __fsqrt_rn, ...), which these flags do not touch.
NVFLAGS = ARCH + ["-lineinfo", "-O3", "-std=c++17", "--expt-relaxed-constexpr", "--extended-lambda", "-prec-div=false", "-prec-sqrt=false",
                  "-Xcompiler", "-fPIC,-ffp-contract=off,-fvisibility=hidden", "-Xptxas", "-v",
                  "-I", os.path.join(ROOT, "include"), "-I", os.path.join(HERE, "csrc")]
CU_SRCS = ["api.cu", "bvh_build.cu", "post_passes.cu"]
HOST_SRCS = ["scene.cpp", "passes.cpp", "image.cpp", "gltf.cpp", "host_c_api.cpp", "multi_device.cpp"]


def run(cmd, log=None):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log:
        open(log, "w").write(r.stdout)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout[-8000:] + "\n")
        raise RuntimeError("build step failed: " + " ".join(cmd[-3:]))
    return r.stdout


def newer(dst, srcs):
    if not os.path.exists(dst):
        return False
    t = os.path.getmtime(dst)
    return all(os.path.getmtime(s) <= t for s in srcs if os.path.exists(s))


def deps(dirs):
    out = []
    for d in dirs:
        for f in os.listdir(d):
            if f.endswith((".h", ".cuh", ".cu", ".cpp")):
                out.append(os.path.join(d, f))
    return out


def build_variant(name, extra_flags):
    """Tuning aid: builds kiraray_b200/lib/libkrr_wfpt_<name>.so with extra nvcc flags (e.g. -DKRR_...=..).
    Select it at run time with KRR_WFPT_LIB=<path> (kiraray_b200/binding.py)."""
    obj = os.path.join(OBJ, name)
    os.makedirs(obj, exist_ok=True)

    def cu(src):
        o = os.path.join(obj, src.replace(".cu", ".o"))
        run([NVCC] + NVFLAGS + list(extra_flags) + ["-c", os.path.join(HERE, "csrc", src), "-o", o], log=os.path.join(obj, src + ".ptxas.log"))
        return o

    with ThreadPoolExecutor(max_workers=4) as ex:
        objs = list(ex.map(cu, CU_SRCS))
    out = os.path.join(LIB, f"libkrr_wfpt_{name}.so")
    run([NVCC] + ARCH + ["-shared", "-cudart", "static", "-Xcompiler", "-fPIC", "-o", out] + objs)
    return out


def check_data():
    """The colour-space tables are an input file (data/MANIFEST.json: provenance, layout, checksum).  Raises when the
    file is missing or does not match the manifest: the host layer would otherwise render with other colours."""
    import hashlib
    import json
    man = json.load(open(os.path.join(HERE, "data", "MANIFEST.json")))
    for name, m in man.items():
        p = os.path.join(HERE, "data", name)
        if not os.path.exists(p):
            raise RuntimeError(f"{p} is missing: {m['how_to_get']}")
        h = hashlib.sha256(open(p, "rb").read()).hexdigest()
        if os.path.getsize(p) != m["bytes"] or h != m["sha256"]:
            raise RuntimeError(f"{p}: size/sha256 {os.path.getsize(p)}/{h} differ from data/MANIFEST.json ({m['bytes']}/{m['sha256']})")
    return True


def build(force=False):
    os.makedirs(OBJ, exist_ok=True)
    hdrs = deps([os.path.join(HERE, "csrc"), os.path.join(HERE, "host"), os.path.join(ROOT, "include")])
    hdrs.append(os.path.abspath(__file__))

    def cu(src):
        s = os.path.join(HERE, "csrc", src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        if force or not newer(o, hdrs):
            run([NVCC] + NVFLAGS + ["-c", s, "-o", o], log=os.path.join(OBJ, src + ".ptxas.log"))
        return o

    with ThreadPoolExecutor(max_workers=4) as ex:
        objs = list(ex.map(cu, CU_SRCS))
    wfpt = os.path.join(LIB, "libkrr_wfpt.so")
    if force or not newer(wfpt, objs):
        run([NVCC] + ARCH + ["-shared", "-cudart", "static", "-Xcompiler", "-fPIC", "-o", wfpt] + objs)
    # exported symbols: the C ABI only (visibility hidden + extern "C" default) -> mark explicitly
    host = os.path.join(LIB, "libkrr_host.so")
    hsrcs = [os.path.join(HERE, "host", s) for s in HOST_SRCS]
    if force or not newer(host, hsrcs + hdrs + [wfpt]):
        run([CXX, "-std=c++17", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-I", os.path.join(ROOT, "include"),
             "-I", "/usr/local/cuda/include", "-o", host] + hsrcs +
            ["-L", LIB, "-lkrr_wfpt", "-L", "/usr/local/cuda/lib64", "-lcudart_static", "-ldl", "-lrt", "-lz", "-pthread",
             "-Wl,-rpath,$ORIGIN"])
    cli_src = os.path.join(HERE, "host", "krr_render.cpp")
    cli = os.path.join(LIB, "krr_render")
    if force or not newer(cli, [cli_src, host]):
        run([CXX, "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), "-o", cli, cli_src,
             "-L", LIB, "-lkrr_host", "-lkrr_wfpt", "-Wl,-rpath,$ORIGIN"])
    return wfpt, host


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "variant":
        print("built", build_variant(sys.argv[2], sys.argv[3:]))
    else:
        build(force="--force" in sys.argv)
        print("built", LIB)
