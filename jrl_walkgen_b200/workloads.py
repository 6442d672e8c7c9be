"""Synthetic workloads of the shapes BASELINE.json names (host-side numpy; no algorithmic product code).

Config 2 - "Kajita2003 preview control batched: 4096 random straight/circle footstep walks, 320-tap preview":
each walk is a footstep list (straight: 8-20 steps of 0.05-0.25 m; circle: arc of radius 0.5-2 m walked with
0.15 m steps, in the spirit of StepStackHandler::CreateArcInStepStack, src/StepStackHandler.cpp:299-457) turned
into a 5 ms ZMP reference with the timing of the reference's ZMPDiscretization (lead-in of 2 preview windows =
640 samples, 160 samples per step = 0.78 s single support + 0.02 s double support, tail of 2 + 960 samples;
src/ZMPRefTrajectoryGeneration/ZMPDiscretization.cpp:319-513, 573-1020, 1129-1300).  The ZMP sits under the
support foot during single support and moves linearly between the feet during double support.  This generator
produces the *shape* of those references for throughput measurements; it is not a parity restatement of
ZMPDiscretization.
"""
from __future__ import annotations

import numpy as np

NL = 320
LEAD_IN = 2 * NL
PER_STEP = 160
SS_SAMPLES = 156
TAIL = 2 + 3 * NL
HALF_FEET = 0.095


def _footsteps(rng):
    """-> (k, 2) support-foot positions of one walk, starting with the right foot."""
    if rng.random() < 0.5:
        n = int(rng.integers(8, 21))
        adv = rng.uniform(0.05, 0.25, size=n)
        heading = np.zeros(n)
    else:
        R = rng.uniform(0.5, 2.0)
        arc = np.deg2rad(rng.uniform(30.0, 180.0))
        n = int(np.clip(np.ceil(arc * R / 0.15), 8, 20))
        adv = np.full(n, arc * R / n)
        heading = np.cumsum(np.full(n, arc / n))
    cx = np.cumsum(adv * np.cos(heading))
    cy = np.cumsum(adv * np.sin(heading))
    side = np.where(np.arange(n) % 2 == 0, -1.0, 1.0) * HALF_FEET
    px = cx - side * np.sin(heading)
    py = cy + side * np.cos(heading)
    return np.stack([px, py], axis=1), (cx[-1], cy[-1])


def preview_walk(rng):
    """One ZMP reference, shape (L, 2) with L = 640 + 160*steps + 962."""
    feet, last = _footsteps(rng)
    n = len(feet)
    L = LEAD_IN + PER_STEP * n + TAIL
    z = np.empty((L, 2))
    z[:LEAD_IN] = 0.0
    prev = np.zeros(2)
    ramp = (np.arange(1, PER_STEP - SS_SAMPLES + 1) / (PER_STEP - SS_SAMPLES + 1))[:, None]
    k = LEAD_IN
    for s in range(n):
        nd = PER_STEP - SS_SAMPLES
        z[k:k + nd] = prev + ramp * (feet[s] - prev)
        z[k + nd:k + PER_STEP] = feet[s]
        prev = feet[s]
        k += PER_STEP
    z[k:] = np.array(last)
    return z


def preview_batch(B, seed=0):
    """Config 2: -> (offsets int64[B+1], zmpref float64[total, 2]); walk b uses stream seed + b."""
    walks = [preview_walk(np.random.default_rng([seed, b])) for b in range(B)]
    lens = np.array([len(w) for w in walks], dtype=np.int64)
    offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    return offsets, np.ascontiguousarray(np.concatenate(walks, axis=0))
