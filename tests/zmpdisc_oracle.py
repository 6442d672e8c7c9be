"""ctypes access to the ZMPDiscretization oracle (oracle/oracle_zmpdisc.cpp) + the step lists of the reference's
TestKajita2003 profiles.  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

REL_STEP_DTYPE = np.dtype([("sx", "f8"), ("sy", "f8"), ("theta", "f8"), ("ss_time", "f8"), ("ds_time", "f8"),
                           ("step_type", "i4"), ("reserved", "i4")])
assert REL_STEP_DTYPE.itemsize == 48


class ZmpDiscParams(C.Structure):
    _fields_ = [("sampling_period", C.c_double), ("preview_time", C.c_double), ("t_single", C.c_double),
                ("t_double", C.c_double), ("step_height", C.c_double), ("omega", C.c_double),
                ("modulation", C.c_double), ("zmp_neutral", C.c_double * 2), ("zmp_shift", C.c_double * 4),
                ("foot_b", C.c_double), ("foot_h", C.c_double), ("foot_f", C.c_double), ("filter_time", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
        _lib.oracle_zmpdisc_run.restype = C.c_long
    return _lib


def default_params():
    p = ZmpDiscParams()
    lib().oracle_zmpdisc_default_params(C.byref(p))
    return p


# initial feet of the HRP-2 half-sitting pose as the reference evaluates them (first row of every Kajita datref):
# left (x, y, theta), right (x, y, theta)
INIT_FEET = np.array([0.00949035, 0.095, 0.0, 0.00949035, -0.095, 0.0])


def steps_from_seq(text, ss=0.78, ds=0.02):
    """:stepseq, walk mode 0 (StepStackHandler.cpp:128-175): triples sx sy theta, SStime/DStime = current defaults."""
    v = np.array(text.split(), dtype=np.float64)
    n = len(v) // 3
    s = np.zeros(n, dtype=REL_STEP_DTYPE)
    s["sx"], s["sy"], s["theta"] = v[0:3 * n:3], v[1:3 * n:3], v[2:3 * n:3]
    s["ss_time"], s["ds_time"], s["step_type"] = ss, ds, 1
    return s


def steps_circle(ss=0.78, ds=0.02, builder=None):
    """TestKajita2003::TurningOnTheCircle (tests/TestKajita2003.cpp:68-92):
    :supportfoot 1, :arc 0.0 0.75 30.0 -1, :lastsupport, :finish."""
    L = builder or lib()
    pre = "oracle_" if builder is None else "wg_"
    s = np.zeros(64, dtype=REL_STEP_DTYPE)
    n = C.c_int(0)
    keep = C.c_int(0)
    p = C.c_void_p(s.ctypes.data)
    assert getattr(L, pre + "steps_support_foot")(p, 64, C.byref(n), 1, C.c_double(ss), C.c_double(ds)) == 0
    assert getattr(L, pre + "steps_arc")(p, 64, C.byref(n), C.c_double(0.0), C.c_double(0.75), C.c_double(30.0), -1,
                                         C.c_double(ss), C.c_double(ds), C.byref(keep)) == 0
    assert getattr(L, pre + "steps_last_support")(p, 64, C.byref(n), keep, C.c_double(ss), C.c_double(ds)) == 0
    return s[:n.value].copy()


STRAIGHT = "0.0 -0.105 0.0 " + "0.2 0.21 0.0 0.2 -0.21 0.0 " * 7 + "0.0 0.21 0.0"

PB_FLORENT_SEQ1 = """0 0.1 0
 -0.0398822 -0.232351 4.6646  -0.0261703 0.199677 4.6646  -0.0471999 -0.256672 4.6646  -0.0305785 0.200634 4.6646
 -0.0507024 -0.245393 4.6646  -0.0339626 0.197227 4.6646  -0.0527259 -0.228579 4.6646  -0.0362332 0.199282 4.6646
 -0.0540087 -0.21638 4.6646  -0.0373302 0.196611 4.6646  -0.0536928 -0.199019 4.6646  -0.0372245 0.204021 4.6646
 -0.0529848 -0.196642 4.6646  -0.0355124 0.2163 4.6646  -0.000858977 -0.204807 0.0767924  0 0.2 0"""

PB_FLORENT_SEQ2 = """0 -0.1 0
 -0.0512076 0.207328 -1.15414  -0.0473172 -0.218623 -1.15414  -0.0515644 0.21034 -1.15414  -0.0475332 -0.215615 -1.15414
 -0.0516395 0.203345 -1.15414  -0.0476688 -0.217615 -1.15414  -0.0517348 0.201344 -1.15414  -0.0477237 -0.219617 -1.15414
 -0.0517494 0.21934 -1.15414  -0.047698 -0.201621 -1.15414  -0.0516832 0.217337 -1.15414  -0.0475915 -0.203622 -1.15414
 -0.0515365 0.215339 -1.15414  -0.0474046 -0.205617 -1.15414  -0.0513094 0.213348 -1.15414  -0.0471374 -0.207603 -1.15414
 -0.0510024 0.216368 -1.15414  -0.0466898 -0.214575 -1.15414  -0.0506158 0.214402 -1.15414  -0.0462637 -0.216533 -1.15414
 -0.0501503 0.217453 -1.15414  -0.0456584 -0.223471 -1.15414  -0.0366673 0.212742 1.62857  -0.0360079 -0.201543 4.21944
 -0.0154622 0.279811 4.21944  -0.0300936 -0.217751 4.21944  -0.00928506 0.283157 4.21944  -0.0236871 -0.219869 4.21944
 -0.00231593 0.290546 4.21944  -0.0169269 -0.202959 4.21944  0.00493436 0.296941 4.21944  -0.00995958 -0.202061 4.21944
 0.0119489 0.297324 4.21944  -0.00293587 -0.202195 4.21944  0.0189437 0.296673 4.21944  0.00399208 -0.203358 4.21944
 0.0257673 0.295003 4.21944  0.0106743 -0.200526 4.21944  0.031904 0.287363 4.21944  0.016966 -0.203651 4.21944
 0.0379487 0.283784 4.21944  0.141438 -0.212069 3.6834  0.204562 0.216453 2.64204  0.200635 -0.218747 -0.366254
 0.216228 0.204108 -2.13008  0.206583 -0.212425 -3.87382  0.187966 0.211947 -6.61811  0.219749 -0.17341 -12.4824
 0.146814 0.240465 -26.5643  0.247166 -0.119114 -37.1489  0.163211 0.222722 -19.7198  0.208825 -0.213706 -5.15242
 0.0285368 0.200005 -0.0337318  0 -0.2 0"""


def profile_steps(name):
    """The four profiles of tests/TestKajita2003.cpp:68-244."""
    if name == "StraightWalking":
        return steps_from_seq(STRAIGHT)
    if name == "Circle":
        return steps_circle()
    if name == "PbFlorentSeq1":
        return steps_from_seq(PB_FLORENT_SEQ1)
    if name == "PbFlorentSeq2":
        return steps_from_seq(PB_FLORENT_SEQ2)
    raise KeyError(name)


def sample_count(p, steps):
    """2*NL + sum_{i>=1} round((DS+SS)/T) + round(Tdble/(2T)) + 3*NL (ZMPDiscretization.cpp:383-386, :638-639,
    :1147-1148, :1239-1241)."""
    T = p.sampling_period
    n = int(2 * p.preview_time / T) + int(round(p.t_double / (2 * T))) + int(3.0 * p.preview_time / T)
    for s in steps[1:]:
        ds, ss = (s["ds_time"], s["ss_time"]) if s["ds_time"] != 0.0 else (p.t_double, p.t_single)
        n += int(round((ds + ss) / T))
    return n


def run(p, steps, init_feet=INIT_FEET):
    """-> dict(zmp [L][3], left [L][6], right [L][6], types [L][3])."""
    steps = np.ascontiguousarray(steps)
    cap = sample_count(p, steps) + 16
    zmp = np.zeros((cap, 3)); left = np.zeros((cap, 6)); right = np.zeros((cap, 6)); types = np.zeros((cap, 3), dtype=np.int32)
    feet = np.ascontiguousarray(init_feet, dtype=np.float64)
    L = lib().oracle_zmpdisc_run(C.byref(p), len(steps), C.c_void_p(steps.ctypes.data), C.c_void_p(feet.ctypes.data),
                                 C.c_long(cap), C.c_void_p(zmp.ctypes.data), C.c_void_p(left.ctypes.data),
                                 C.c_void_p(right.ctypes.data), C.c_void_p(types.ctypes.data))
    if L < 0:
        raise RuntimeError(f"oracle_zmpdisc_run failed: {L}")
    return {"zmp": zmp[:L], "left": left[:L], "right": right[:L], "types": types[:L]}


def golden(name):
    g = np.load(os.path.join(ROOT, "tests", "golden", "kajita_%s.npz" % name))
    return g["q"].astype(np.float64) / float(g["scale"])
