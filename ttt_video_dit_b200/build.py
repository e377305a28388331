"""Build libttt_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python ttt_video_dit_b200/build.py [--force] [--verbose]
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libttt_b200.so")
SELFTEST_LIB = os.path.join(LIBDIR, "libttt_b200_selftest.so")  # include/ttt_b200_debug.h: development probes only
SELFTEST_SRC = ("umma_selftest.cu", "dsmem_probe.cu", "capi_debug.cu", "tmap.cu")   # tmap.cu is shared with the production library
PROD_EXCLUDE = ("umma_selftest.cu", "dsmem_probe.cu", "capi_debug.cu")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC",
         "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def needs_build():
    if not os.path.exists(LIB) or not os.path.exists(SELFTEST_LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", h) for h in ("ttt_b200.h", "ttt_b200_debug.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, dbg=False, noelect=False):
    global LIB
    if noelect:  # A/B variant of the single-lane issue pattern (csrc/ptx.cuh TB_NO_ELECT)
        return _build(os.path.join(LIBDIR, "libttt_b200_noelect.so"), os.path.join(HERE, "build_noelect"), ["-DTB_NO_ELECT"], verbose)
    if dbg:
        return _build(os.path.join(LIBDIR, "libttt_b200_dbg.so"), os.path.join(HERE, "build_dbg"), ["-DTTT_PHASE_TIMING"], verbose)
    if not force and not needs_build():
        return LIB
    return _build(LIB, os.path.join(HERE, "build"), [], verbose)


def _build(LIB, objdir, extra, verbose):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(objdir, exist_ok=True)
    procs = []
    for s in sources():
        o = os.path.join(objdir, s[:-3] + ".o")
        cmd = [NVCC, *FLAGS, *extra, "-c", os.path.join(CSRC, s), "-o", o]
        procs.append((s, o, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    objs = []
    log = []
    for s, o, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {s}\n{out}")
        if p.returncode != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError(f"nvcc failed on {s}")
        objs.append(o)
    arch = ["-gencode", "arch=compute_100a,code=sm_100a"]
    obj_of = lambda s: os.path.join(objdir, s[:-3] + ".o")
    subprocess.check_call([NVCC, "-shared", "-o", LIB, *[o for o in objs if os.path.basename(o) not in
                                                          [x[:-3] + ".o" for x in PROD_EXCLUDE]], *arch])
    if not extra:  # release build: the self-test library next to it
        subprocess.check_call([NVCC, "-shared", "-o", SELFTEST_LIB, *[obj_of(s) for s in SELFTEST_SRC], *arch])
    with open(os.path.join(objdir, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv, dbg="--dbg" in sys.argv, noelect="--noelect" in sys.argv))
