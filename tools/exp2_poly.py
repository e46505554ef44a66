import numpy as np
from numpy.polynomial import chebyshev as C
def fit(deg, lo=-0.5, hi=0.5, iters=30):
    # Remez-like iteration for relative error of 2^x: minimise max |p(x)/2^x - 1|
    # weighted least squares on dense grid with iterative reweighting (Lawson)
    x = np.linspace(lo, hi, 20001)
    y = 2.0 ** x
    w = np.ones_like(x)
    V = np.vander(x, deg + 1, increasing=True)
    for _ in range(200):
        A = V * (w / y)[:, None]
        b = w
        c, *_ = np.linalg.lstsq(A, b, rcond=None)
        e = np.abs(V @ c / y - 1)
        w = w * (e / e.max()) ** 0.5 + 1e-12
        w /= w.max()
    return c, e.max()
for deg in (2, 3, 4):
    c, e = fit(deg)
    print(deg, e, [float(v) for v in c])
    # float32 emulation
    c32 = c.astype(np.float32)
    xs = np.float32(np.linspace(-40, 0, 2000001, dtype=np.float64))
    magic = np.float32(12582912.0)
    t = (xs + magic).astype(np.float32)
    n = (t - magic).astype(np.float32)
    f = (xs - n).astype(np.float32)
    p = np.full_like(f, c32[deg])
    for k in range(deg - 1, -1, -1):
        p = (p * f + c32[k]).astype(np.float32)   # not fused, close enough
    bits = p.view(np.int32) + (t.view(np.int32) << 23)
    r = bits.view(np.float32)
    ref = 2.0 ** xs.astype(np.float64)
    rel = np.abs(r.astype(np.float64) / ref - 1)
    print("  f32 emu max rel err", rel.max(), "at x", xs[rel.argmax()])
