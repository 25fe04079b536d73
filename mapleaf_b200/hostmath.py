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
