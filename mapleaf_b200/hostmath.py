"""
Scalar vector / quaternion helpers for host-side model construction (initial state, CG shift).

Same operation order as the reference's Cython value types so that host-computed constants are
bit-identical: Motion/CythonVector.pyx:70-139, Motion/CythonQuaternion.pyx:37-56,102-169.
Vectors are 3-tuples, quaternions 4-tuples (Q0 scalar first).
"""
from math import acos, cos, sin, sqrt


def v_add(a, b):
    return (a[0] + b[0], a[1] + b[1], a[2] + b[2])


def v_sub(a, b):
    # reference: self + (-b)
    return (a[0] + -b[0], a[1] + -b[1], a[2] + -b[2])


def v_scale(a, s):
    return (a[0] * s, a[1] * s, a[2] * s)


def v_div(a, s):
    # reference: self * (1/scalar)
    r = 1 / s
    return (a[0] * r, a[1] * r, a[2] * r)


def v_dot(a, b):
    return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]


def v_cross(a, b):
    return (a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0])


def v_len(a):
    return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2])


def v_normalize(a):
    l = v_len(a)
    if l > 0:
        return (a[0] / l, a[1] / l, a[2] / l)
    return (0.0, 0.0, 0.0)


def v_angle(a, b):
    interior = v_dot(a, b) / (v_len(a) * v_len(b))
    if abs(interior - 1) < 0.00000001:
        return acos(1)
    return acos(interior)


def q_from_axis_angle(axis, angle):
    ax = v_normalize(axis)
    return (cos(angle / 2), sin(angle / 2) * ax[0], sin(angle / 2) * ax[1], sin(angle / 2) * ax[2])


def q_mul(p, q):
    return (p[0] * q[0] - p[1] * q[1] - p[2] * q[2] - p[3] * q[3],
            p[0] * q[1] + p[1] * q[0] + p[2] * q[3] - p[3] * q[2],
            p[0] * q[2] - p[1] * q[3] + p[2] * q[0] + p[3] * q[1],
            p[0] * q[3] + p[1] * q[2] - p[2] * q[1] + p[3] * q[0])


def q_norm(q):
    return sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3])


def q_normalize(q):
    n = q_norm(q)
    if n == 1.0:
        return q
    return (q[0] / n, q[1] / n, q[2] / n, q[3] / n)


def q_conj(q):
    return (q[0], -q[1], -q[2], -q[3])


def q_rotate(q, v):
    """Returns (rotated vector, normalized quaternion): the reference normalizes the receiver in place."""
    rq = q_normalize(q)
    t = q_mul(q_mul(rq, (0, v[0], v[1], v[2])), q_conj(rq))
    return (t[1], t[2], t[3]), rq


# ---------------------------------------------------------------------------------------------
# Earth models: launch-tower-frame state -> global inertial frame (ENV/EarthModelling.py:107-227,242-341)
# ---------------------------------------------------------------------------------------------
from math import atan2, degrees, pi, radians  # noqa: E402

EARTH_ROTATION_RATE = 7.292115e-5
SPHERE_RADIUS = 6371007.1809
SPHERE_GM = 3.98600436e14
WGS_A = 6378137
WGS_F = 1 / 298.257223563
WGS_B = WGS_A * (1 - WGS_F)
WGS_E2 = WGS_F * (2 - WGS_F)


def geodeticToCartesian(model, lat, lon, height):
    lat = radians(lat)
    lon = radians(lon)
    lon += EARTH_ROTATION_RATE * 0
    if model == "Round":
        radius = SPHERE_RADIUS + height
        return (radius * cos(lat) * cos(lon), radius * cos(lat) * sin(lon), radius * sin(lat))
    N = WGS_A / sqrt(1 - WGS_E2 * sin(lat) ** 2)
    return ((N + height) * cos(lat) * cos(lon), (N + height) * cos(lat) * sin(lon), ((WGS_B ** 2 / WGS_A ** 2) * N + height) * sin(lat))


def cartesianToGeodetic(model, x, y, z):
    if model == "Round":
        p = sqrt(x * x + y * y)
        lon = atan2(y, x)
        lon -= EARTH_ROTATION_RATE * 0
        lat = atan2(z, p)
        height = sqrt(x * x + y * y + z * z) - SPHERE_RADIUS
        return degrees(lat), degrees(lon), height
    lon = atan2(y, x)
    lon -= EARTH_ROTATION_RATE * 0
    p = sqrt(x * x + y * y)
    Kold = 1 / (1 - WGS_E2)
    Korig = Kold
    while True:
        ci = ((p * p + (1 - WGS_E2) * z * z * Kold ** 2) ** (1.5)) / (WGS_A * WGS_E2)
        Knew = 1 + (p * p + (1 - WGS_E2) * z * z * Kold ** 3) / (ci - p * p)
        if abs(Knew - Kold) < 1e-10:
            break
        Kold = Knew
    k = Knew
    lat = atan2(k * z, p)
    h = (1 / WGS_E2) * (1 / k - 1 / Korig) * sqrt(p * p + z * z * k * k)
    return degrees(lat), degrees(lon), h


def inertialToENURotation(model, x, y, z):
    """SphericalEarth.getInertialToENUFrameRotation (:170-190), inherited by WGS84 with its own cartesianToGeodetic."""
    lat, lon, _ = cartesianToGeodetic(model, x, y, z)
    rot1 = q_from_axis_angle((0, 0, 1), radians(lon + 90))
    rot2 = q_from_axis_angle((1, 0, 0), radians(90 - lat))
    return q_mul(rot1, rot2)


def convertIntoGlobalFrame(model, pos, vel, q, w, lat, lon):
    """SphericalEarth.convertIntoGlobalFrame (:197-227); pos.z is already ASL."""
    gpos = geodeticToCartesian(model, lat, lon, pos[2])
    towerToGlobal = inertialToENURotation(model, *gpos)
    rotatedVelocity, towerToGlobal = q_rotate(towerToGlobal, vel)
    earthAngVel = (0, 0, EARTH_ROTATION_RATE)
    gvel = v_add(rotatedVelocity, v_cross(earthAngVel, gpos))
    gq = q_mul(towerToGlobal, q)
    angVelRocketFrame, _ = q_rotate(q_conj(gq), earthAngVel)
    gw = v_add(w, angVelRocketFrame)
    return gpos, gvel, gq, gw
