"""Extract golden vectors from the reference's own test data into small committed fixtures.

Run in the build container (needs /root/reference); the GPU box only sees the committed .npz files.
  herdt_online_prefix.npz : tests/TestHerdt2010OnLineTestFGPI.datref.cmake rows 0..4999 (t < 25 s: the
                            translation-only prefix; later rows involve robot-specific hip-yaw limits that are
                            not in the container), stored as int64 of value*1e7 (the datref is truncated
                            to 7 decimals by TestObject.cpp:48-56).
  herdt_emergency_prefix.npz : tests/TestHerdt2010EmergencyStopTestFGPI.datref.cmake rows 0..1027.
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
    a = np.loadtxt(os.path.join(REF, "tests", "TestHerdt2010OnLineTestFGPI.datref.cmake"))
    np.savez_compressed(os.path.join(HERE, "herdt_online_prefix.npz"), q=quantised(a[:5000]),
                        source="tests/TestHerdt2010OnLineTestFGPI.datref.cmake rows 0..4999", scale=1e7)
    b = np.loadtxt(os.path.join(REF, "tests", "TestHerdt2010EmergencyStopTestFGPI.datref.cmake"))
    np.savez_compressed(os.path.join(HERE, "herdt_emergency_prefix.npz"), q=quantised(b[:1028]),
                        source="tests/TestHerdt2010EmergencyStopTestFGPI.datref.cmake rows 0..1027", scale=1e7)
    cols = [10, 11, 12, 19, 20, 21, 22, 23, 24, 31, 32, 33, 34, 35]
    for prof in ("StraightWalking", "Circle", "PbFlorentSeq1", "PbFlorentSeq2"):
        src = "tests/TestKajita2003%sTestFGPI.datref.cmake" % prof
        k = np.loadtxt(os.path.join(REF, src))
        np.savez_compressed(os.path.join(HERE, "kajita_%s.npz" % prof), q=quantised(k[:, cols]), cols=np.array(cols) + 1,
                            t0=k[0, 0], source=src, scale=1e7)
    print("written", [f for f in os.listdir(HERE) if f.endswith(".npz")])


if __name__ == "__main__":
    main()
