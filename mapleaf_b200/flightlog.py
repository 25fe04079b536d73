"""
Per-step trajectory logs of selected lanes in the reference's own file format.

With SimControl.loggingLevel >= 1 the reference keeps a TimeStepLog per rocket (Rocket/rocket.py:155-170: position, velocity,
orientation quaternion, NED Tait-Bryan Euler angles, angular velocity, and the adaptive integrator's error estimate) with one row per
Rocket.timeStep call (`_logState`, :615-644) plus the final state, and writes it as `<top stage>_timeStepLog.csv`
(IO/CythonLog.pyx:163-213, rocket.py:729-738). The kernels' flight log (mlb_set_flight_log with every = 1: one 16-double row per accepted
step, written by the lane itself: 128 B of HBM writes per step) carries the same information; this module turns a lane's rows into that
file, column for column:
  * columns sorted by name after `Time(s)`; vectors expanded to X / Y / Z, the quaternion to its first THREE components 0 / 1 / 2
    (expandIterableColumns loops over range(3) whatever the length);
  * `TimeStep(s)` = differences of the times, the last one repeated (TimeStepLog._calculateTimeStepSizes);
  * `EstimatedIntegrationError` (adaptive methods only) is logged into the row of the state the step STARTS from; the final row gets the
    fill value 0;
  * rows logged under a parachute (3-DoF body) have no orientation / angular velocity / Euler angles: fill values, printed 0.0;
  * the logged quaternion is the state's object, which the following step normalises in place: all rows but the last show it normalised.
Values are written with repr(), as csv.writer does for Python floats.
"""
import csv
import math
from typing import Sequence

import numpy as np

from . import hostmath as hm
from .model import L

__all__ = ["timeStepLogColumns", "writeTimeStepLog"]


def _eulerAnglesNED(model, pos, q):
    """Quaternion.toEulerAngles of the orientation relative to the local NED frame (rocket.py:630-636, EarthModelling.py:33-36,
    CythonQuaternion.pyx:211-217)."""
    earth = model.meta["earth"]
    toENU = hm.inertialToENURotation(earth, *pos) if earth in ("Round", "WGS84") else (1.0, 0.0, 0.0, 0.0)
    ned = hm.q_mul(toENU, hm.q_from_axis_angle((1, 1, 0), math.pi))
    q0, q1, q2, q3 = hm.q_mul(hm.q_conj(ned), q)
    return (math.atan2(2 * (q0 * q1 + q2 * q3), 1 - 2 * (q1 ** 2 + q2 ** 2)), math.asin(2 * (q0 * q2 - q3 * q1)),
            math.atan2(2 * (q0 * q3 + q1 * q2), 1 - 2 * (q2 ** 2 + q3 ** 2)))


def timeStepLogColumns(rows: np.ndarray, model) -> dict:
    """{column name: list of values} of the reference's time-step log for one lane's flight-log rows ([n, 16]: time, position, velocity,
    quaternion, angular velocity, step size taken, error estimate; row 0 = the initial state)."""
    rows = np.asarray(rows, dtype=np.float64)
    n = rows.shape[0]
    adaptive = "Adapt" in model.meta["method"]
    # the body turns 3-DoF at the first parachute deployment and stays so: from then on the log rows hold the identity orientation
    sixDof = np.ones(n, dtype=bool)
    if model.header("H_RECOVERY_OFF") != 0:
        isIdentity = (rows[:, 7] == 1.0) & (rows[:, 8:14] == 0.0).all(axis=1)
        k = n
        while k > 1 and isIdentity[k - 1]:
            k -= 1
        if k < n and k > 0:
            # the deployment happens at the START of the step that produced row k (event detector before `_logState`, rocket.py:669-687), so
            # the state that step started from, row k - 1, was already logged as a 3-DoF state
            sixDof[k - 1:] = False
    cols = {"Time(s)": [0 if (i == 0 and rows[0, 0] == 0) else float(rows[i, 0]) for i in range(n)]}
    for name, c0 in (("Position", 1), ("Velocity", 4)):
        unit = "(m)" if name == "Position" else "(m/s)"
        for j, ax in enumerate("XYZ"):
            cols["{}{}{}".format(name, ax, unit)] = [float(x) for x in rows[:, c0 + j]]
    # the log holds the state's quaternion OBJECT, and the step that starts from a state normalises that object in place
    # (RigidBodyStates.py:110-117 -> CythonQuaternion.pyx:144-152): every logged orientation but the last is the normalised one, while the
    # Euler angles were computed from it as it was at logging time
    quats = []
    for i in range(n):
        q = [float(x) for x in rows[i, 7:11]]
        if i < n - 1 and sixDof[i]:
            nrm = math.sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3])
            if nrm != 1.0:
                q = [c / nrm for c in q]
        quats.append(q)
    for j in range(3):
        cols["OrientationQuaternion{}".format(j)] = [quats[i][j] if sixDof[i] else 0.0 for i in range(n)]
        cols["AngularVelocity{}(rad/s)".format("XYZ"[j])] = [float(rows[i, 11 + j]) if sixDof[i] else 0.0 for i in range(n)]
    euler = [_eulerAnglesNED(model, rows[i, 1:4], tuple(rows[i, 7:11])) if sixDof[i] else (0.0, 0.0, 0.0) for i in range(n)]
    for j, ax in enumerate("XYZ"):
        cols["EulerAngle{}(rad)".format(ax)] = [e[j] for e in euler]
    times = rows[:, 0]
    if n > 1:
        steps = [float(times[i + 1] - times[i]) for i in range(n - 1)]
        cols["TimeStep(s)"] = steps + [steps[-1]]
    else:
        cols["TimeStep(s)"] = []
    if adaptive:
        cols["EstimatedIntegrationError"] = [float(rows[i + 1, 15]) for i in range(n - 1)] + [0]
    return cols


def writeTimeStepLog(path: str, rows: Sequence, model) -> str:
    """Writes `<top stage>_timeStepLog.csv` as the reference would for this flight (IO/CythonLog.pyx:163-190)."""
    cols = timeStepLogColumns(rows, model)
    names = sorted(cols.keys())
    names.remove("Time(s)")
    names.insert(0, "Time(s)")
    n = len(cols["Time(s)"])
    with open(path, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(names)
        for i in range(n):
            w.writerow([cols[c][i] for c in names])
    return path
