"""ctypes binding of ``libcsg2im.so`` (the C ABI declared in ``include/csg2im.h``).

The header is the single source of truth: prototypes are parsed from it, so a
symbol that is declared but not exported (or the other way round) fails at load
time.  There is no fallback of any kind: if the library is missing or fails to
load, importing the kernels raises.
"""
import ctypes
import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
HEADER = os.path.join(ROOT, "include", "csg2im.h")
LIB_PATH = os.path.join(PKG, "libcsg2im.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]

_C2CT = {
    "int": ctypes.c_int,
    "long long": ctypes.c_longlong,
    "size_t": ctypes.c_size_t,
    "double": ctypes.c_double,
    "void": None,
    "csg_stream_t": ctypes.c_void_p,
}


def parse_header(path=HEADER):
    """Return {name: (restype, [argtypes])} for every prototype in the header."""
    src = open(path).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(csg_\w+)\s*\(([^;{}]*?)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if "typedef" in ret:
            continue

        def conv(t):
            t = t.strip()
            if "*" in t:
                return ctypes.c_char_p if t.replace(" ", "") == "constchar*" else ctypes.c_void_p
            t = t.replace("const", "").strip()
            return _C2CT[t]

        restype = conv(ret)
        argtypes = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                # drop the parameter name (last identifier) unless the token is a bare type
                mm = re.match(r"(.*?)(\b\w+)$", a, flags=re.S)
                ty = mm.group(1).strip() if mm and mm.group(1).strip() else a
                argtypes.append(conv(ty))
        protos[name] = (restype, argtypes)
    return protos


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ for sm_100a and link libcsg2im.so in-tree."""
    if not force and not needs_build():
        return LIB_PATH
    objdir = os.path.join(PKG, "build")
    os.makedirs(objdir, exist_ok=True)
    common = [NVCC] + ARCH_FLAGS + ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-fvisibility=hidden",
                                    "-I", os.path.join(ROOT, "include"), "-I", CSRC]

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        cmd = common + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose and (r.stdout or r.stderr):
            print(r.stdout, r.stderr, file=sys.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(compile_one, sources()))
    tmp = LIB_PATH + ".tmp"
    link = [NVCC] + ARCH_FLAGS + ["-shared", "-o", tmp] + objs + ["-cudart", "static"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


_lib = None


def load():
    """Load the shared library and attach the header's prototypes.  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libcsg2im.so is not built (%s missing). Run `python -c 'import __graft_entry__ as g; g.build()'`; "
            "there is no CPU/PyTorch fallback for the csg2im kernels." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in parse_header().items():
        fn = getattr(lib, name)            # AttributeError = declared in the header but not exported
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


class CsgError(RuntimeError):
    pass


_ASYNC_WHAT = {1: "triple subject/object id", 2: "triple predicate id", 3: "embedding id", 4: "canonicalization triplet"}
_async_rec = (ctypes.c_int * 4)()


def poll_async_errors(synchronize=False):
    """Raise ``IndexError`` if a kernel met an out-of-range index (include/csg2im.h, csg_async_error_poll).

    The reference raises IndexError synchronously on such inputs (graph.py:63-64,73,98-103, model.py:108-109);
    kernels cannot, so they neutralise the row and leave a record that the next library call (every ``check``)
    or an explicit ``poll_async_errors(synchronize=True)`` turns into the exception."""
    if synchronize:
        import torch
        torch.cuda.synchronize()
    code = load().csg_async_error_poll(_async_rec)
    if code:
        _, row, value, limit = list(_async_rec)
        raise IndexError("csg2im: %s out of range: value %d at row %d is not in [0, %d)"
                         % (_ASYNC_WHAT.get(code, "index (code %d)" % code), value, row, limit))


def check(rc, what=""):
    if rc != 0:
        msg = load().csg_last_error()
        raise CsgError("%s failed (rc=%d): %s" % (what or "csg call", rc, msg.decode() if msg else "?"))
    poll_async_errors()
