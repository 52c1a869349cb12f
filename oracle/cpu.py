"""TEST INFRASTRUCTURE ONLY.  ctypes bindings to

* ``oracle/liboracle.so``          -- the plain-C restatement (d3feat_oracle.c), kind "port"
* ``oracle/_ref/libd3feat_ref.so`` -- the UNMODIFIED reference C++ behind ref_shim.cpp, kind "reference"

Both expose the two native entry points the reference dataloader calls
(datasets/dataloader.py:17,63): ``subsample_batch`` and ``batch_query``.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_F = ctypes.POINTER(ctypes.c_float)
_I = ctypes.POINTER(ctypes.c_int)


def build(verbose=False):
    """Compile liboracle.so (always) and oracle/_ref (only where /root/reference exists)."""
    out = subprocess.run(["make", "-C", _HERE, "all"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if verbose:
        print(out.stdout)


def _load(path):
    if not os.path.exists(path):
        build()
    if not os.path.exists(path):
        return None
    return ctypes.CDLL(path)


_port = None
_ref = None


def port_lib():
    global _port
    if _port is None:
        _port = _load(os.path.join(_HERE, "liboracle.so"))
        if _port is None:
            raise RuntimeError("oracle/liboracle.so missing and could not be built")
        _port.orc_radius_neighbors.restype = ctypes.c_int
        _port.orc_radius_neighbors.argtypes = [_F, ctypes.c_int, _F, ctypes.c_int, _I, _I, ctypes.c_int,
                                               ctypes.c_float, ctypes.POINTER(_I), _I]
        _port.orc_grid_subsampling.restype = ctypes.c_int
        _port.orc_grid_subsampling.argtypes = [_F, ctypes.c_int, _I, ctypes.c_int, ctypes.c_float, _F, _I]
        _port.orc_free.argtypes = [ctypes.c_void_p]
    return _port


def ref_available():
    return os.path.exists(os.path.join(_HERE, "_ref", "libd3feat_ref.so")) or os.path.isdir("/root/reference")


def ref_lib():
    global _ref
    if _ref is None:
        _ref = _load(os.path.join(_HERE, "_ref", "libd3feat_ref.so"))
        if _ref is None:
            raise RuntimeError("oracle/_ref/libd3feat_ref.so missing (reference tree absent and no prebuilt copy)")
        _ref.ref_batch_neighbors.restype = _I
        _ref.ref_batch_neighbors.argtypes = [_F, ctypes.c_int, _F, ctypes.c_int, _I, _I, ctypes.c_int,
                                             ctypes.c_float, _I]
        _ref.ref_batch_grid_subsampling.restype = ctypes.c_int
        _ref.ref_batch_grid_subsampling.argtypes = [_F, ctypes.c_int, _I, ctypes.c_int, ctypes.c_float,
                                                    ctypes.c_int, _F, _I]
        _ref.ref_free.argtypes = [ctypes.c_void_p]
    return _ref


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def batch_query(queries, supports, q_batches, s_batches, radius, impl="port"):
    """radius_neighbors.batch_query (cpp_neighbors/wrapper.cpp:58) -> int32 [Nq, max_count]."""
    q, s, ql, sl = _f32(queries), _f32(supports), _i32(q_batches), _i32(s_batches)
    nq, ns, nb = q.shape[0], s.shape[0], ql.shape[0]
    mc = ctypes.c_int(0)
    if impl == "port":
        lib = port_lib()
        ptr = _I()
        rc = lib.orc_radius_neighbors(q.ctypes.data_as(_F), nq, s.ctypes.data_as(_F), ns,
                                      ql.ctypes.data_as(_I), sl.ctypes.data_as(_I), nb,
                                      ctypes.c_float(radius), ctypes.byref(ptr), ctypes.byref(mc))
        assert rc == 0
        free = lib.orc_free
    else:
        lib = ref_lib()
        ptr = lib.ref_batch_neighbors(q.ctypes.data_as(_F), nq, s.ctypes.data_as(_F), ns,
                                      ql.ctypes.data_as(_I), sl.ctypes.data_as(_I), nb,
                                      ctypes.c_float(radius), ctypes.byref(mc))
        free = lib.ref_free
    n = nq * mc.value
    out = np.ctypeslib.as_array(ptr, shape=(max(n, 1),))[:n].copy().reshape(nq, mc.value)
    free(ptr)
    return out


def subsample_batch(points, batches, sampleDl, max_p=0, impl="port"):
    """grid_subsampling.subsample_batch (cpp_subsampling/wrapper.cpp:62), points only -> (s_points, s_len)."""
    p, bl = _f32(points), _i32(batches)
    n, nb = p.shape[0], bl.shape[0]
    out = np.empty((max(n, 1), 3), np.float32)
    out_len = np.zeros(nb, np.int32)
    if impl == "port":
        assert max_p == 0
        m = port_lib().orc_grid_subsampling(p.ctypes.data_as(_F), n, bl.ctypes.data_as(_I), nb,
                                            ctypes.c_float(sampleDl), out.ctypes.data_as(_F),
                                            out_len.ctypes.data_as(_I))
    else:
        m = ref_lib().ref_batch_grid_subsampling(p.ctypes.data_as(_F), n, bl.ctypes.data_as(_I), nb,
                                                 ctypes.c_float(sampleDl), int(max_p),
                                                 out.ctypes.data_as(_F), out_len.ctypes.data_as(_I))
    return out[:m].copy(), out_len
