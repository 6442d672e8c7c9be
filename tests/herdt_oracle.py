"""ctypes access to the Herdt part of the oracle (oracle/oracle_herdt.cpp).  TEST INFRASTRUCTURE."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N = 16
MAX_VARS = 36
MAX_ROWS = 75


from jrl_walkgen_b200._capi import HerdtParams, herdt_dtypes  # plain struct mirrors (no compute)

QP_INPUT_DTYPE, QP_OUTPUT_DTYPE = herdt_dtypes()

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
        L.oracle_herdt_set_ref_lib.argtypes = [C.c_char_p]
        L.oracle_herdt_have_ref_qld.restype = C.c_int
        L.oracle_herdt_default_params.argtypes = [C.c_double, C.c_double, C.POINTER(HerdtParams)]
        L.oracle_herdt_build_qp.argtypes = [C.POINTER(HerdtParams), C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int),
                                            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_herdt_solve_qp.argtypes = [C.POINTER(HerdtParams), C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_herdt_solve_qp.restype = C.c_int
        L.oracle_herdt_solve_qp_batch.argtypes = [C.POINTER(HerdtParams), C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_herdt_solve_qp_batch.restype = C.c_long
        L.oracle_herdt_sim_new.restype = C.c_void_p
        L.oracle_herdt_sim_new.argtypes = [C.POINTER(HerdtParams), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_int, C.c_int]
        L.oracle_herdt_sim_delete.argtypes = [C.c_void_p]
        L.oracle_herdt_sim_set_robot.argtypes = [C.c_void_p] + [C.c_double] * 5
        L.oracle_herdt_sim_set_return_to_centre.argtypes = [C.c_void_p, C.c_int]
        L.oracle_herdt_sim_set_initial_support.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        L.oracle_herdt_sim_set_vel_ref.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        L.oracle_herdt_sim_steps_before_stop.argtypes = [C.c_void_p, C.c_uint]
        L.oracle_herdt_sim_stoppg.argtypes = [C.c_void_p]
        L.oracle_herdt_sim_tick.argtypes = [C.c_void_p, C.c_void_p]
        L.oracle_herdt_sim_tick.restype = C.c_int
        L.oracle_herdt_sim_num_qp.argtypes = [C.c_void_p]
        L.oracle_herdt_sim_last_fail.argtypes = [C.c_void_p]
        L.oracle_herdt_sim_get_log.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.oracle_herdt_sim_get_log.restype = C.c_int
        ref = os.path.join(ROOT, "oracle", "_ref", "libwalkgen_ref.so")
        if os.path.exists(ref):
            L.oracle_herdt_set_ref_lib(ref.encode())
        _lib = L
    return _lib


def default_params(sole_length=0.25, sole_width=0.14):
    """Sole 0.25 x 0.14 m: fitted from the reference's datref."""
    p = HerdtParams()
    lib().oracle_herdt_default_params(sole_length, sole_width, C.byref(p))
    return p


def load_golden(name):
    """tests/golden/herdt_<name>_full.npz -> the reference datref as float rows [time, 37 columns] (see make_golden.py)."""
    z = np.load(os.path.join(ROOT, "tests", "golden", "herdt_%s_full.npz" % name))
    q = np.concatenate([z["q0"][None, :].astype(np.int64), z["dq"].astype(np.int64)], axis=0).cumsum(axis=0)
    return q / 1e7


class Sim:
    """Closed-loop TestHerdt2010 harness around the oracle."""

    def __init__(self, params=None, com0=(0.0316055, 0.0, 0.7116911), lf0=(0.0, 0.09, 0.0), rf0=(0.0, -0.09, 0.0),
                 zmp0=(0.0, 0.0, 0.0), textbook=False, logging=False):
        self.p = params or default_params()
        a = [np.array(v, dtype=np.float64) for v in (com0, lf0, rf0, zmp0)]
        self.h = lib().oracle_herdt_sim_new(C.byref(self.p), *[v.ctypes.data for v in a], int(textbook), int(logging))
        self.row = np.zeros(37)

    def initial_support(self, x, y, yaw):
        lib().oracle_herdt_sim_set_initial_support(self.h, x, y, yaw)

    def datref_era(self):
        """The three settings that make the surveyed code path reproduce the reference's committed datrefs, which are
        older than the source (oracle pins): (1) initial double-support frame at (0, 0.1, 0) (pre-3.1.8
        InitOnLine), (2) hip-yaw velocity bound 0 = what OrientationsPreview.cpp:67 reads from a robot file without
        velocity limits (the test robot's; hip angle limits equal -> the -30/+45 deg defaults of :52-53), (3) no
        return-to-centre jerk at the end of the walk (ZMPVelocityReferencedQP.cpp:410-421 was added in 3.1.8)."""
        self.initial_support(0.0, 0.1, 0.0)
        d = np.pi / 180.0
        lib().oracle_herdt_sim_set_robot(self.h, -30 * d, 45 * d, -30 * d, 45 * d, 0.0)
        lib().oracle_herdt_sim_set_return_to_centre(self.h, 0)

    def vel_ref(self, x, y, yaw):
        lib().oracle_herdt_sim_set_vel_ref(self.h, x, y, yaw)

    def steps_before_stop(self, n):
        lib().oracle_herdt_sim_steps_before_stop(self.h, n)

    def stoppg(self):
        lib().oracle_herdt_sim_stoppg(self.h)

    def tick(self):
        rc = lib().oracle_herdt_sim_tick(self.h, self.row.ctypes.data)
        return rc, self.row.copy()

    def log(self):
        n = lib().oracle_herdt_sim_num_qp(self.h)
        ins = np.zeros(n, dtype=QP_INPUT_DTYPE)
        x = np.zeros((n, MAX_VARS)); u = np.zeros((n, MAX_ROWS + 1)); meta = np.zeros((n, 3), dtype=np.int32)
        k = lib().oracle_herdt_sim_get_log(self.h, ins.ctypes.data, x.ctypes.data, u.ctypes.data, meta.ctypes.data, n)
        return ins[:k], x[:k], u[:k], meta[:k]

    def close(self):
        if self.h:
            lib().oracle_herdt_sim_delete(self.h)
            self.h = None


# tests/TestHerdt2010.cpp:225-244 (generateEventOnLineWalking) and :258-262 (generateEventEmergencyStop):
# (tick at which the ParseCmd is issued, :setVelReference arguments, followed by :stoppg?)
ONLINE_SCRIPT = [(5 * 200, (0.2, 0.0, 0.0), False),           # walkForward
                 (10 * 200, (0.0, 0.2, 0.0), False),          # walkSidewards
                 (25 * 200, (0.0, 0.0, -10.0), False),        # startTurningRightOnSpot
                 (35 * 200, (0.2, 0.0, 0.0), False),
                 (45 * 200, (0.0, 0.0, 10.0), False),         # startTurningLeftOnSpot
                 (55 * 200, (0.2, 0.0, 0.0), False),
                 (65 * 200, (0.0, 0.0, -10.0), False),
                 (75 * 200, (0.2, 0.0, 0.0), False),
                 (85 * 200, (0.2, 0.0, 6.0832), False),       # startTurningLeft
                 (95 * 200, (0.2, 0.0, -6.0832), False),      # startTurningRight
                 (105 * 200, (0.0, 0.0, 0.0), False),         # stop
                 (110 * 200, (0.0, 0.0, 0.0), True)]          # stopOnLineWalking (:setVelReference 0 0 0 + :stoppg)
EMERGENCY_SCRIPT = [(5 * 200, (0.0, 0.0, 0.4), False),        # startTurningLeft2
                    (10 * 200, (0.2, 0.0, -0.2), False),      # startTurningRight2
                    (3040, (0.0, 0.0, 0.0), False),           # 15.2*200 stop
                    (4160, (0.0, 0.0, 0.0), True)]            # 20.8*200 stopOnLineWalking


def events_from(script):
    def make(v, stop):
        def ev(s):
            s.vel_ref(*v)
            if stop:
                s.stoppg()
        return ev
    return {t: make(v, stop) for t, v, stop in script}


ONLINE_EVENTS = events_from(ONLINE_SCRIPT)
EMERGENCY_EVENTS = events_from(EMERGENCY_SCRIPT)


def run_online_script(nticks, events, initial_support=None, datref_era=False, **kw):
    """TestHerdt2010 OnLine: tests/TestHerdt2010.cpp:66-90 (start) and :232-244 (events).  `events` maps the tick
    index (m_OneStep.NbOfIt at the time generateEvent() runs) to a callable(sim)."""
    sim = Sim(**kw)
    sim.steps_before_stop(2)
    if datref_era:
        sim.datref_era()
    if initial_support is not None:
        sim.initial_support(*initial_support)
    rows = np.zeros((nticks, 37))
    for it in range(nticks):
        rc, row = sim.tick()
        rows[it] = row
        if it in events:
            events[it](sim)
    return sim, rows
