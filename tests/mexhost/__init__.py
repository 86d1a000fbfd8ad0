"""Host emulator for matlab/epi_mex.cpp: builds the gateway against tests/mexhost/mex_emul.cpp (an in-memory
implementation of the MEX C API subset of tests/stubs/mex.h) and calls mexFunction out of process, so that the
gateway's marshalling code is EXECUTED -- on the GPU box against the real library -- without MATLAB/Octave.

    out = call("ekf_eks", 0, u, x, params_dict, ..., nlhs=11)       # numpy in, numpy out (MATLAB shapes)

dict -> 1x1 struct, list of dicts -> 1xn struct array, str -> char row, bool arrays -> logical,
everything else -> double array (column-major, at least 2-D like MATLAB).  Test infrastructure."""
import os
import struct
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
BIN = os.path.join(HERE, "_build", "mex_driver")


class MexError(RuntimeError):
    def __init__(self, ident, msg):
        super().__init__(f"{ident}: {msg}")
        self.id, self.msg = ident, msg


def build(force=False):
    """g++ mex_emul.cpp mex_driver.cpp matlab/epi_mex.cpp -lepi_b200 (rpath to the in-tree library)."""
    libdir = os.path.join(ROOT, "epidemicmodeling_b200")
    srcs = [os.path.join(HERE, "mex_emul.cpp"), os.path.join(HERE, "mex_driver.cpp"),
            os.path.join(ROOT, "matlab", "epi_mex.cpp")]
    deps = srcs + [os.path.join(HERE, "mex_emul.h"), os.path.join(ROOT, "tests", "stubs", "mex.h"),
                   os.path.join(ROOT, "include", "epi_b200.h"), os.path.join(libdir, "libepi_b200.so")]
    if not force and os.path.exists(BIN) and all(os.path.getmtime(BIN) >= os.path.getmtime(d) for d in deps):
        return BIN
    os.makedirs(os.path.dirname(BIN), exist_ok=True)
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", HERE, "-I", os.path.join(ROOT, "tests", "stubs"),
           "-I", os.path.join(ROOT, "include")] + srcs + ["-L", libdir, "-lepi_b200", f"-Wl,-rpath,{libdir}", "-o", BIN]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("mex host build failed:\n" + r.stderr)
    return BIN


def _node(v):
    if v is None:
        return struct.pack("<i", 4)
    if isinstance(v, str):
        b = v.encode()
        return struct.pack("<iiqq", 1, 2, 1, len(b)) + b
    if isinstance(v, dict) or (isinstance(v, (list, tuple)) and v and isinstance(v[0], dict)):
        elems = [v] if isinstance(v, dict) else list(v)
        names = list(elems[0].keys())
        out = struct.pack("<iiqq", 2, 2, 1, len(elems)) + struct.pack("<i", len(names))
        for n in names:
            out += struct.pack("<i", len(n)) + n.encode()
        for e in elems:
            for n in names:
                out += _node(e.get(n))
        return out
    a = np.asarray(v)
    if a.dtype == np.bool_:
        a2 = np.atleast_2d(a)
        return struct.pack("<ii", 3, a2.ndim) + struct.pack(f"<{a2.ndim}q", *a2.shape) + \
            np.asfortranarray(a2).astype(np.uint8).tobytes(order="F")
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 0:
        a = a.reshape(1, 1)
    elif a.ndim == 1:
        a = a.reshape(1, -1)              # MATLAB vectors from Python 1-D arrays are rows
    return struct.pack("<ii", 0, a.ndim) + struct.pack(f"<{a.ndim}q", *a.shape) + a.tobytes(order="F")


def _read(buf, pos):
    tag, = struct.unpack_from("<i", buf, pos)
    pos += 4
    if tag == 4:
        return None, pos
    nd, = struct.unpack_from("<i", buf, pos)
    pos += 4
    dims = struct.unpack_from(f"<{nd}q", buf, pos)
    pos += 8 * nd
    n = int(np.prod(dims))
    if tag == 0:
        a = np.frombuffer(buf, dtype="<f8", count=n, offset=pos).reshape(dims, order="F").copy()
        return a, pos + 8 * n
    if tag == 1:
        return bytes(buf[pos:pos + n]).decode(), pos + n
    if tag == 3:
        a = np.frombuffer(buf, dtype=np.uint8, count=n, offset=pos).reshape(dims, order="F").astype(bool)
        return a, pos + n
    raise ValueError("unexpected node in reply")


def call(cmd, *args, nlhs=1, timeout=300):
    """epi_mex(cmd, args...) with nlhs outputs; raises MexError for mexErrMsgIdAndTxt."""
    exe = build()
    with tempfile.TemporaryDirectory() as d:
        fi, fo = os.path.join(d, "in.bin"), os.path.join(d, "out.bin")
        with open(fi, "wb") as fh:
            fh.write(struct.pack("<ii", nlhs, 1 + len(args)))
            fh.write(_node(cmd))
            for a in args:
                fh.write(_node(a))
        r = subprocess.run([exe, fi, fo], capture_output=True, text=True, timeout=timeout)
        if r.returncode != 0:
            raise RuntimeError(f"mex_driver exited {r.returncode}: {r.stderr}")
        buf = open(fo, "rb").read()
    n, = struct.unpack_from("<i", buf, 0)
    pos = 4
    if n < 0:
        l, = struct.unpack_from("<i", buf, pos); pos += 4
        ident = buf[pos:pos + l].decode(); pos += l
        l, = struct.unpack_from("<i", buf, pos); pos += 4
        raise MexError(ident, buf[pos:pos + l].decode())
    outs = []
    for _ in range(n):
        v, pos = _read(buf, pos)
        outs.append(v)
    return outs
