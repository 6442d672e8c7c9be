"""Extract golden vectors from the reference's own test data into small committed fixtures.

Run in the build container (needs /root/reference); the GPU box only sees the committed .npz files.
  herdt_online_full.npz   : tests/TestHerdt2010OnLineTestFGPI.datref.cmake, all 22 348 rows x 38 columns, stored as
                            int64 of value*1e7 (the datref is truncated to 7 decimals by TestObject.cpp:48-56):
                            first row `q0` + row-to-row differences `dq` (int32; compresses 10x better).
                            Read back with tests/herdt_oracle.py:load_golden().
  herdt_emergency_full.npz : tests/TestHerdt2010EmergencyStopTestFGPI.datref.cmake, all 4 508 rows, same coding.
  kajita_<profile>.npz    : tests/TestKajita2003<profile>TestFGPI.datref.cmake, every row, columns 11-13, 20-22
                            (left foot x y z theta omega omega2), 23-25, 32-34 (right foot), 35-36 (world ZMP
                            reference): the outputs of ZMPDiscretization + FootTrajectoryGenerationStandard, which do
                            not depend on the proprietary HRP-2 model.  Also column 1 of the first row.
"""
import os
import sys

import numpy as np

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def quantised(a):
    q = np.rint(a * 1e7).astype(np.int64)
    assert np.abs(q / 1e7 - a).max() < 1e-12, "datref has more than 7 decimals?"
    return q


def main():
    for name, out in (("TestHerdt2010OnLineTestFGPI", "herdt_online_full.npz"),
                      ("TestHerdt2010EmergencyStopTestFGPI", "herdt_emergency_full.npz")):
        src = "tests/%s.datref.cmake" % name
        q = quantised(np.loadtxt(os.path.join(REF, src)))
        dq = np.diff(q, axis=0)
        assert np.abs(dq).max() < 2**31
        np.savez_compressed(os.path.join(HERE, out), q0=q[0], dq=dq.astype(np.int32), source=src, scale=1e7)
    cols = [10, 11, 12, 19, 20, 21, 22, 23, 24, 31, 32, 33, 34, 35]
    for prof in ("StraightWalking", "Circle", "PbFlorentSeq1", "PbFlorentSeq2"):
        src = "tests/TestKajita2003%sTestFGPI.datref.cmake" % prof
        k = np.loadtxt(os.path.join(REF, src))
        np.savez_compressed(os.path.join(HERE, "kajita_%s.npz" % prof), q=quantised(k[:, cols]), cols=np.array(cols) + 1,
                            t0=k[0, 0], source=src, scale=1e7)
    print("written", [f for f in os.listdir(HERE) if f.endswith(".npz")])


if __name__ == "__main__":
    main()
