"""TEST INFRASTRUCTURE (oracle) - deterministic float32 math used by the simulator spec.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product (copo_b200/) never does.

MetaDrive 0.2.5 (the reference's simulator, README.md:41-42) is not vendored, so there is no
reference arithmetic to follow here; these routines are the repo's own spec for sin/cos/atan2 so
that the numpy oracle and the CUDA kernel (compiled with -fmad=false) agree bit for bit:
every operation below is a single IEEE-754 binary32 add / mul / div / sqrt, in the order written.
"""
import numpy as np

f32 = np.float32
u32 = np.uint32


def _c(x):
    return np.float32(x)


# Cody-Waite split of pi/2 (cephes DP1..3 doubled) and cephes sinf/cosf minimax coefficients.
PIO2_HI = _c(1.5703125)
PIO2_MID = _c(4.837512969970703125e-4)
PIO2_LO = _c(7.549789948768648e-8)
TWO_OVER_PI = _c(0.6366197723675814)
SIN_C1 = _c(-1.6666654611e-1)
SIN_C2 = _c(8.3321608736e-3)
SIN_C3 = _c(-1.9515295891e-4)
COS_C1 = _c(4.166664568298827e-2)
COS_C2 = _c(-1.388731625493765e-3)
COS_C3 = _c(2.443315711809948e-5)
PI = _c(3.14159265358979323846)
TWO_PI = _c(6.28318530717958647692)
HALF_PI = _c(1.57079632679489661923)
# Abramowitz & Stegun 4.4.49: atan(a)/a on [0,1], |err| <= 2e-8
ATAN_C = [_c(v) for v in (1.0, -0.3333314528, 0.1999355085, -0.1420889944, 0.1065626393, -0.0752896400,
                          0.0429096138, -0.0161657367, 0.0028662257)]

CONSTS = dict(PIO2_HI=PIO2_HI, PIO2_MID=PIO2_MID, PIO2_LO=PIO2_LO, TWO_OVER_PI=TWO_OVER_PI, SIN_C1=SIN_C1,
              SIN_C2=SIN_C2, SIN_C3=SIN_C3, COS_C1=COS_C1, COS_C2=COS_C2, COS_C3=COS_C3, PI=PI, TWO_PI=TWO_PI,
              HALF_PI=HALF_PI, **{"ATAN_C%d" % i: v for i, v in enumerate(ATAN_C)})


def det_sincos(x):
    """(sin, cos) of float32 array x; valid for |x| < ~100."""
    x = np.asarray(x, dtype=f32)
    k = np.rint(x * TWO_OVER_PI).astype(f32)
    r = x - k * PIO2_HI
    r = r - k * PIO2_MID
    r = r - k * PIO2_LO
    q = k.astype(np.int32) & 3
    r2 = r * r
    p = SIN_C3 * r2
    p = p + SIN_C2
    p = p * r2
    p = p + SIN_C1
    p = p * r2
    p = p * r
    s = r + p
    c = COS_C3 * r2
    c = c + COS_C2
    c = c * r2
    c = c + COS_C1
    c = c * r2
    c = c * r2
    h = _c(0.5) * r2
    c = c - h
    c = c + _c(1.0)
    sin_o = np.where(q == 0, s, np.where(q == 1, c, np.where(q == 2, -s, -c)))
    cos_o = np.where(q == 0, c, np.where(q == 1, -s, np.where(q == 2, -c, s)))
    return sin_o.astype(f32), cos_o.astype(f32)


def det_atan2(y, x):
    y = np.asarray(y, dtype=f32)
    x = np.asarray(x, dtype=f32)
    ax = np.abs(x)
    ay = np.abs(y)
    swap = ay > ax
    mx = np.where(swap, ay, ax)
    mn = np.where(swap, ax, ay)
    with np.errstate(divide="ignore", invalid="ignore"):
        a = np.where(mx == _c(0.0), _c(0.0), mn / np.where(mx == _c(0.0), _c(1.0), mx)).astype(f32)
    s = a * a
    p = ATAN_C[8]
    for i in range(7, -1, -1):
        p = p * s
        p = p + ATAN_C[i]
    r = a * p
    r = np.where(swap, HALF_PI - r, r)
    r = np.where(x < _c(0.0), PI - r, r)
    r = np.where(y < _c(0.0), -r, r)
    return r.astype(f32)


def wrap_pi(h):
    h = np.where(h > PI, h - TWO_PI, h)
    h = np.where(h < -PI, h + TWO_PI, h)
    return h.astype(f32)


# --- counter-based RNG (lowbias32 finaliser) --------------------------------------------------------
def _mix(x):
    x = x.astype(np.uint64)
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x7FEB352D)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(15)
    x = (x * np.uint64(0x846CA68B)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    return x


def rng_u32(seed, scene, episode, ctr):
    """u32 draw keyed by (seed, scene, episode, counter); all integer arrays broadcast."""
    m = np.uint64(0xFFFFFFFF)
    seed = np.asarray(seed).astype(np.uint64) & m
    scene = np.asarray(scene).astype(np.uint64) & m
    episode = np.asarray(episode).astype(np.uint64) & m
    ctr = np.asarray(ctr).astype(np.uint64) & m
    x = _mix(seed ^ ((scene * np.uint64(0x9E3779B1)) & m))
    x = _mix(x ^ ((episode * np.uint64(0x85EBCA77)) & m))
    x = _mix(x ^ ((ctr * np.uint64(0xC2B2AE3D)) & m))
    return x.astype(np.uint32)


def u32_to_unit(u):
    """Top 24 bits -> float32 in [0, 1) (exact)."""
    return ((np.asarray(u, dtype=np.uint32) >> np.uint32(8)).astype(f32) * _c(1.0 / 16777216.0)).astype(f32)
