"""ctypes access to the reference's OWN PreviewControl / OptimalControllerSolver object code
(oracle/_ref/libwalkgen_ref.so, built by oracle/Makefile from /root/reference/src/PreviewControl/*.cpp where they
lie; glue: oracle/ref_glue_preview.cc).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import glob
import os
import sysconfig

import numpy as np

import oracle_lib as ol

D = ol.D
_bound = False
_lapack = None


def lib():
    global _bound
    r = ol.ref()
    if r is None or not hasattr(r, "ref_preview_new"):
        return None
    if not _bound:
        r.ref_preview_new.restype = C.c_void_p
        r.ref_preview_new.argtypes = [C.c_uint, C.c_int]
        r.ref_preview_delete.argtypes = [C.c_void_p]
        r.ref_preview_read_file.argtypes = [C.c_void_p, C.c_char_p]
        r.ref_preview_compute_weights.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_uint]
        r.ref_preview_call_method.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        r.ref_preview_get_gains.restype = C.c_int
        r.ref_preview_get_gains.argtypes = [C.c_void_p, D, D, D, D, D, D, C.c_int, D]
        r.ref_preview_set_gains.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double, D, C.c_double, D, C.c_int]
        r.ref_preview_run.restype = C.c_long
        r.ref_preview_run.argtypes = [C.c_void_p, D, C.c_long, D, D, D, C.c_int, C.c_int]
        r.ref_preview_run_batch.restype = C.c_long
        r.ref_preview_run_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, ol.I64, D, D, D, D, C.c_int]
        r.ref_preview_run_1d_deque.restype = C.c_long
        r.ref_preview_run_1d_deque.argtypes = [C.c_void_p, D, C.c_long, D, D, D, C.c_int]
        r.ref_preview_step_1d_vector.restype = C.c_int
        r.ref_preview_step_1d_vector.argtypes = [C.c_void_p, D, C.c_long, C.c_uint, D, D, D, C.c_int]
        r.ref_lapack_open.restype = C.c_int
        r.ref_lapack_open.argtypes = [C.c_char_p]
        _bound = True
    return r


def lapack_available():
    """Open the OpenBLAS that ships inside the SciPy (or OpenCV) wheel of this image: the reference expects a system
    LAPACK for dgges_ (OptimalControllerSolver.cpp:44-52) and the image has none."""
    global _lapack
    if _lapack is not None:
        return _lapack
    r = lib()
    _lapack = False
    if r is None:
        return False
    site = sysconfig.get_paths()["purelib"]
    cands = sorted(glob.glob(os.path.join(site, "scipy.libs", "libscipy_openblas-*.so"))) + \
        sorted(glob.glob(os.path.join(site, "opencv_python_headless.libs", "libopenblasp-*.so")))
    for p in cands:
        if r.ref_lapack_open(p.encode()) == 0:
            _lapack = True
            break
    return _lapack


class RefPreview:
    """One PreviewControl object of the reference."""

    def __init__(self, mode=1, auto=False):
        self.r = lib()
        self.h = self.r.ref_preview_new(mode, int(auto))

    def close(self):
        if self.h:
            self.r.ref_preview_delete(self.h)
            self.h = None

    def compute_weights(self, T, preview_time, zc, mode):
        self.r.ref_preview_compute_weights(self.h, T, preview_time, zc, mode)

    def call_method(self, method, args):
        self.r.ref_preview_call_method(self.h, method.encode(), args.encode())

    def read_file(self, path):
        self.r.ref_preview_read_file(self.h, path.encode())

    def gains(self):
        A = np.zeros(9); B = np.zeros(3); Cm = np.zeros(3); Kx = np.zeros(3); F = np.zeros(4096); par = np.zeros(3)
        ks = C.c_double()
        nl = self.r.ref_preview_get_gains(self.h, ol.dptr(A), ol.dptr(B), ol.dptr(Cm), ol.dptr(Kx), C.byref(ks), ol.dptr(F),
                                          4096, ol.dptr(par))
        return dict(A=A, B=B, C=Cm, Kx=Kx, Ks=ks.value, F=F[:nl].copy(), NL=nl, T=par[0], preview_time=par[1], zc=par[2])

    def set_gains(self, T, preview_time, zc, Kx, Ks, F):
        Kx = np.ascontiguousarray(Kx, dtype=np.float64); F = np.ascontiguousarray(F, dtype=np.float64)
        self.r.ref_preview_set_gains(self.h, T, preview_time, zc, ol.dptr(Kx), Ks, ol.dptr(F), len(F))

    def run(self, zmpref_xy, state8, simulation=True, use_lindex=False):
        """OneIterationOfPreview over a whole trajectory; returns (com [steps][6], zmp [steps][2], steps)."""
        z = np.ascontiguousarray(zmpref_xy, dtype=np.float64)
        L = len(z)
        com = np.zeros((max(L, 1), 6)); zmp = np.zeros((max(L, 1), 2))
        steps = self.r.ref_preview_run(self.h, ol.dptr(z), L, ol.dptr(state8), ol.dptr(com), ol.dptr(zmp), int(simulation),
                                       int(use_lindex))
        return com, zmp, steps

    def run_batch(self, offsets, zmpref_xy, state, com, zmp, b0=0, b1=None, stride=1, simulation=True):
        """Walks b0, b0+stride, ... < b1 of a ragged batch; outputs in place.  Returns the number of preview steps."""
        b1 = len(offsets) - 1 if b1 is None else b1
        return self.r.ref_preview_run_batch(self.h, b0, b1, stride, offsets.ctypes.data_as(ol.I64), ol.dptr(zmpref_xy),
                                            ol.dptr(state), ol.dptr(com), ol.dptr(zmp), int(simulation))

    def run_1d_deque(self, zmpref, state4, simulation=True):
        z = np.ascontiguousarray(zmpref, dtype=np.float64)
        com = np.zeros((max(len(z), 1), 3)); zmp = np.zeros(max(len(z), 1))
        steps = self.r.ref_preview_run_1d_deque(self.h, ol.dptr(z), len(z), ol.dptr(state4), ol.dptr(com), ol.dptr(zmp),
                                                int(simulation))
        return com, zmp, steps

    def step_1d_vector(self, zmpref, lindex, x3, sxzmp, simulation=True):
        z = np.ascontiguousarray(zmpref, dtype=np.float64)
        x = np.array(x3, dtype=np.float64); s = C.c_double(sxzmp); zo = C.c_double()
        rc = self.r.ref_preview_step_1d_vector(self.h, ol.dptr(z), len(z), lindex, ol.dptr(x), C.byref(s), C.byref(zo),
                                               int(simulation))
        return rc, x, s.value, zo.value


def write_precomputed_file(path, zc, T, preview_time, Kx, Ks, F, digits=17):
    """The file format ReadPrecomputedFile parses (PreviewControl.cpp:142-176): zc T Tprev Kx[3] Ks F[NL]."""
    fmt = "%." + str(digits) + "g"
    with open(path, "w") as f:
        f.write(" ".join(fmt % v for v in (zc, T, preview_time)) + "\n")
        f.write(" ".join(fmt % v for v in Kx) + "\n")
        f.write(fmt % Ks + "\n")
        f.write("\n".join(fmt % v for v in F) + "\n")
