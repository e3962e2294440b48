#!/usr/bin/env python3
"""Build recipe for the CPU oracle (TEST INFRASTRUCTURE ONLY -- the product never links this).

  python oracle/build_oracle.py ref      -> oracle/_ref/libkrr_oracle_ref.so     (reference's own
                                            KRR_CALLABLE headers, compiled where they lie under
                                            /root/reference/src; needs that tree, so it can only be
                                            built in the dev container -- the prebuilt .so travels)
  python oracle/build_oracle.py spectral -> kiraray_b200/data/spectral_srgb.bin  (colour-space tables
                                            the reference host application owns, dumped via the ref
                                            backend; they are INPUT DATA of the C ABI)

Nothing from /root/reference is copied into git-tracked paths: patched shadows of the four
reference files that need a one-line change to compile with g++ are GENERATED into oracle/_ref/gen/
(git-ignored) by the edits listed in `SHADOW_EDITS` below (SURVEY.md section 8c lists the blockers).
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("KRR_REFERENCE_ROOT", "/root/reference")
REF_OUT = os.path.join(HERE, "_ref")
GEN = os.path.join(REF_OUT, "gen")
CXX = os.environ.get("CXX", "g++")
# -ffp-contract=off: no FMA contraction, so float results do not depend on -march
COMMON = ["-std=c++17", "-fPIC", "-pthread", "-ffp-contract=off", "-fno-fast-math"]


def run(cmd):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout[-6000:] + "\n")
        raise RuntimeError("oracle build step failed: " + cmd[-1])
    return r.stdout


def newer(dst, srcs):
    if not os.path.exists(dst):
        return False
    t = os.path.getmtime(dst)
    return all(os.path.exists(s) and os.path.getmtime(s) <= t for s in srcs)


# ------------------------------------------------------------------------------------------------
# reference build
# ------------------------------------------------------------------------------------------------
# (reference-relative source, generated destination, [(old, new), ...] or callable)
def _slice_hg(text):
    """media.cpp: keep the includes and the host-compilable HG phase-function definitions
    (reference src/render/media.cpp:83-111); the majorant-grid CUDA kernel above them is dropped."""
    head = text[: text.index("NAMESPACE_BEGIN(krr)")] + "NAMESPACE_BEGIN(krr)\n"
    start = text.index("KRR_HOST_DEVICE PhaseFunctionSample HGPhaseFunction::sample")
    return head + text[start:]


SHADOW_EDITS = [
    # '"..."##LOG' token paste is MSVC-only
    ("src/util/check.h", "shadow/util/check.h", [('"##LOG', '" LOG')]),
    # these include "check.h" relative to their own directory: they must sit beside the shadow
    ("src/util/lowdiscrepancy.h", "shadow/util/lowdiscrepancy.h", []),
    ("src/util/tables.h", "shadow/util/tables.h", []),
    # Eigen::Transform -> krr::Transform needs a converting constructor under g++
    ("src/core/math/include/krrmath/transform.h", "shadow/krrmath/transform.h",
     [("\tusing Eigen::Transform<T, Dim, Mode, Options>::Transform;",
       "\tusing Eigen::Transform<T, Dim, Mode, Options>::Transform;\n"
       "\tKRR_CALLABLE Transform(const Eigen::Transform<T, Dim, Mode, Options>& o) : "
       "Eigen::Transform<T, Dim, Mode, Options>(o) {}")]),
    # copy-init of XYZ from an Eigen product expression
    ("src/render/color.cpp", "src/color.cpp",
     [("XYZ C\t   = rgb.inverse() * Vector3f(W);",
       "Vector3f Cv = rgb.inverse() * Vector3f(W); XYZ C(Cv[0], Cv[1], Cv[2]);")]),
    ("src/core/device/gpustd.cpp", "src/gpustd.cpp",
     [("_aligned_malloc(size, alignment)",
       "aligned_alloc(alignment, (size + alignment - 1) / alignment * alignment)"),
      ("_aligned_free(ptr)", "free(ptr)")]),
    ("src/render/media.cpp", "src/media_hg.cpp", _slice_hg),
]

REF_INCLUDES = [
    "{gen}/shadow", "{here}/ref/compat", "{here}", "{ref}/src", "{ref}/src/core",
    "{ref}/src/core/math/3rdparty/eigen", "{ref}/src/core/math/include", "{ref}/src/ext/json",
    "{ref}/src/ext/openvdb/nanovdb/include", "{ref}/src/ext/glfw/include", "{ref}/src/ext/imgui",
    "{ref}/src/ext/nvrhi/include", "{ref}/src/ext/nvrhi/thirdparty/Vulkan-Headers/include",
    "{ref}/src/ext/image", "/usr/local/cuda/include",
    # generated (patched) sources include siblings of their original directories
    "{ref}/src/core/device", "{ref}/src/render",
]


def gen_shadows():
    kdir = os.path.join(REF, "src/core/math/include/krrmath")
    os.makedirs(os.path.join(GEN, "shadow/krrmath"), exist_ok=True)
    # krrmath headers include each other relatively -> the whole directory must be shadowed
    for f in os.listdir(kdir):
        if f != "transform.h":
            shutil.copyfile(os.path.join(kdir, f), os.path.join(GEN, "shadow/krrmath", f))
    for src, dst, edits in SHADOW_EDITS:
        text = open(os.path.join(REF, src), encoding="utf-8", errors="replace").read()
        if callable(edits):
            text = edits(text)
        else:
            for old, new in edits:
                if old not in text:
                    raise RuntimeError(f"shadow edit no longer applies to {src}: {old!r}")
                text = text.replace(old, new)
        out = os.path.join(GEN, dst)
        os.makedirs(os.path.dirname(out), exist_ok=True)
        open(out, "w").write(text)


def build_ref():
    if not os.path.isdir(os.path.join(REF, "src")):
        print(f"[oracle] {REF} absent: keeping prebuilt oracle/_ref (if any)")
        return os.path.exists(os.path.join(REF_OUT, "libkrr_oracle_ref.so"))
    os.makedirs(os.path.join(GEN, "obj"), exist_ok=True)
    gen_shadows()
    inc = []
    for i in REF_INCLUDES:
        inc += ["-I", i.format(gen=GEN, here=HERE, ref=REF)]
    flags = COMMON + ["-fpermissive", "-w", "-include", os.path.join(HERE, "ref/compat/compat.h")] + inc
    units = [  # (source, opt)
        (os.path.join(REF, "src/render/spectrum.cpp"), "-O1"),
        (os.path.join(GEN, "src/color.cpp"), "-O1"),
        (os.path.join(GEN, "src/gpustd.cpp"), "-O1"),
        (os.path.join(GEN, "src/media_hg.cpp"), "-O2"),
        (os.path.join(REF, "src/util/tables.cpp"), "-O1"),
        (os.path.join(REF, "src/data/rgbspectrum_srgb.cpp"), "-O0"),
        # the three colour spaces the path never uses are zero-filled stand-ins (see the file's header)
        (os.path.join(HERE, "ref/unused_tables.cpp"), "-O0"),
        (os.path.join(HERE, "ref/backend_ref.cpp"), "-O2"),
        (os.path.join(HERE, "driver.cpp"), "-O2"),
    ]
    deps_extra = [os.path.join(HERE, "oracle_leaf.h"), os.path.join(HERE, "driver.h"),
                  os.path.join(ROOT, "include/krr_wfpt.h")]

    def compile_one(u):
        src, opt = u
        obj = os.path.join(GEN, "obj", os.path.basename(src).replace(".cpp", ".o"))
        if newer(obj, [src] + (deps_extra if HERE in src else [])):
            return obj
        run([CXX] + flags + [opt, "-I", os.path.join(ROOT, "include"), "-c", src, "-o", obj])
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        objs = list(ex.map(compile_one, units))
    so = os.path.join(REF_OUT, "libkrr_oracle_ref.so")
    # -Bsymbolic: the fake CUDA runtime inside this .so must win over a real libcudart that the
    # host process (torch) may already have loaded into the global scope
    run([CXX, "-shared", "-pthread", "-Wl,-Bsymbolic", "-o", so] + objs)
    print("[oracle] built", so)
    return True


def dump_spectral():
    """Dump the sRGB colour-space data (CIE X/Y/Z, D65, RGB<->XYZ, RGB->spectrum table) that the
    reference host application owns and passes to every pass (wavefront.h:28 `colorSpace`)."""
    import ctypes
    out = os.path.join(ROOT, "kiraray_b200/data/spectral_srgb.bin")
    so = os.path.join(REF_OUT, "libkrr_oracle_ref.so")
    if os.path.exists(out) and (not os.path.exists(so) or os.path.getmtime(out) >= os.path.getmtime(so)):
        return True
    if not os.path.exists(so):
        print("[oracle] cannot dump spectral tables: reference backend not built")
        return False
    lib = ctypes.CDLL(so)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    rc = lib.ol_ref_dump_spectral(out.encode())
    if rc != 0:
        raise RuntimeError("ol_ref_dump_spectral failed")
    print("[oracle] wrote", out, os.path.getsize(out), "bytes")
    return True


def main(argv):
    what = argv[1:] or ["ref", "spectral"]
    for w in what:
        {"ref": build_ref, "spectral": dump_spectral}[w]()


if __name__ == "__main__":
    main(sys.argv)
