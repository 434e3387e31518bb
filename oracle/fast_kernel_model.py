"""
oracle/fast_kernel_model.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Instruction-level CPU model of the fast RK4 kernel's arithmetic (csrc/yaman4.cu: `stage()` and the
z-loop of `fast_integrate`), with every FMA evaluated exactly and rounded once (via 80-bit long
double).  It exists so that the kernel's ALGORITHM -- constant step, stage states written directly,
RK4 combination rebuilt from the stage states, phase by rotation recurrence with re-sync -- can be
checked against the oracle (the reference's arithmetic) on a CPU, and so that rounding-bias bugs
can be reproduced without a GPU (the 2^-54 weight bias of tests/test_gpu_edges.py was found this
way).  It must be kept in step with yaman4.cu by hand; tests/test_gpu_parity.py ties the real
kernel to it on the GPU.  Pure-Python loops: small cases only.
"""
from __future__ import annotations

import math

import numpy as np

_L = np.longdouble
RESYNC = 32


def fma(a, b, c):
    """a*b + c rounded once to double (exact product and sum fit 64-bit significands closely enough:
    double rounding differs from a true FMA only in rare half-way cases)."""
    return float(_L(a) * _L(b) + _L(c))


def stage(inn, base, qr, qi, cg, c2g, cn):
    """out = base + c*f(in) -- the operation order of `stage()` in csrc/yaman4.cu."""
    x1, y1, x2, y2, x3, y3, x4, y4 = inn
    P1 = fma(y1, y1, x1 * x1)
    P2 = fma(y2, y2, x2 * x2)
    P3 = fma(y3, y3, x3 * x3)
    P4 = fma(y4, y4, x4 * x4)
    S = (P1 + P2) + (P3 + P4)
    c2 = c2g * S
    G1, G2, G3, G4 = (fma(-cg, P, c2) for P in (P1, P2, P3, P4))
    Ur, Ui = fma(-y3, y4, x3 * x4), fma(x3, y4, y3 * x4)
    Vr, Vi = fma(-y1, y2, x1 * x2), fma(x1, y2, y1 * x2)
    Wr, Wi = fma(-qi, Ui, qr * Ur), fma(qr, Ui, qi * Ur)
    Zr, Zi = fma(qi, Vi, qr * Vr), fma(qr, Vi, -(qi * Vr))
    return [
        fma(-G1, y1, fma(y2, Wr, fma(-x2, Wi, fma(cn, x1, base[0])))),
        fma(G1, x1, fma(x2, Wr, fma(y2, Wi, fma(cn, y1, base[1])))),
        fma(-G2, y2, fma(y1, Wr, fma(-x1, Wi, fma(cn, x2, base[2])))),
        fma(G2, x2, fma(x1, Wr, fma(y1, Wi, fma(cn, y2, base[3])))),
        fma(-G3, y3, fma(y4, Zr, fma(-x4, Zi, fma(cn, x3, base[4])))),
        fma(G3, x3, fma(x4, Zr, fma(y4, Zi, fma(cn, y3, base[5])))),
        fma(-G4, y4, fma(y3, Zr, fma(-x3, Zi, fma(cn, x4, base[6])))),
        fma(G4, x4, fma(x3, Zr, fma(y3, Zi, fma(cn, y4, base[7])))),
    ]


def weights(kind="shipped"):
    """(third, two_thirds, third_c): weights of y/ys2, ys3 and ys4 in the rebuilt RK4 combination.
    `shipped`: two_thirds + third_c == 1 exactly.  `naive`: fl(1/3), fl(2/3), fl(1/3) (sum 1 - 2^-54)."""
    third, two_thirds = 1.0 / 3.0, 2.0 / 3.0
    return (third, two_thirds, 1.0 - two_thirds) if kind == "shipped" else (third, two_thirds, third)


def integrate(A0, gamma, alpha, dbeta, z_max, n_steps, *, save_every=None, weight_kind="shipped"):
    """The fast kernel's z-loop for one point; returns the list of saved states (complex[4]) incl. z = 0."""
    h = z_max / n_steps
    w = [0.5 * h, h, h * (1.0 / 6.0)]
    nha = -0.5 * alpha
    cg = [wi * gamma for wi in w]
    c2g = [c + c for c in cg]
    cn = [wi * nha for wi in w]
    q0 = c2g[0]
    third, two_thirds, third_c = weights(weight_kind)
    y = []
    for a in np.asarray(A0, dtype=np.complex128):
        y += [a.real, a.imag]
    rr, ri = math.cos(dbeta * (0.5 * h)), math.sin(dbeta * (0.5 * h))
    qr, qi = q0, 0.0
    saved = [np.array([complex(y[2 * j], y[2 * j + 1]) for j in range(4)])]
    for i in range(n_steps):
        if i % RESYNC == 0:
            ang = dbeta * fma(float(i), h, 0.0)
            qr, qi = q0 * math.cos(ang), q0 * math.sin(ang)
        qhr, qhi = fma(-qi, ri, qr * rr), fma(qr, ri, qi * rr)
        q2r, q2i = qhr + qhr, qhi + qhi
        qfr, qfi = fma(-qhi, ri, qhr * rr), fma(qhr, ri, qhi * rr)
        q6r, q6i = qfr * third, qfi * third
        ys = stage(y, y, qr, qi, cg[0], c2g[0], cn[0])
        acc = [fma(third, ys[j], y[j] * (-third)) for j in range(8)]
        yt = stage(ys, y, qhr, qhi, cg[0], c2g[0], cn[0])
        acc = [fma(two_thirds, yt[j], acc[j]) for j in range(8)]
        ys = stage(yt, y, q2r, q2i, cg[1], c2g[1], cn[1])
        acc = [fma(third_c, ys[j], acc[j]) for j in range(8)]
        y = stage(ys, acc, q6r, q6i, cg[2], c2g[2], cn[2])
        qr, qi = qfr, qfi
        if save_every and (i + 1) % save_every == 0:
            saved.append(np.array([complex(y[2 * j], y[2 * j + 1]) for j in range(4)]))
    if not save_every:
        saved.append(np.array([complex(y[2 * j], y[2 * j + 1]) for j in range(4)]))
    return saved
