"""Scratch: F(4,3) interpolation points and the numerical error of F(2x4,3x3) with the tensor-core accumulation
model of wino_emul.py.  Matrices by Cook-Toom; B^T solved numerically from the identity."""
import numpy as np, itertools, sys
sys.path.insert(0, 'scratch')
from fractions import Fraction as Fr

def cook_toom(points, m, r):
    n = m + r - 1
    a = [Fr(p) for p in points]; assert len(a) == n - 1
    f = [np.prod([a[i] - a[j] for j in range(n - 1) if j != i]) for i in range(n - 1)]
    AT = np.array([[float(a[k] ** i) for k in range(n - 1)] + [1.0 if i == m - 1 else 0.0] for i in range(m)])
    G = np.array([[float(a[k] ** j / f[k]) for j in range(r)] for k in range(n - 1)] + [[0.0] * (r - 1) + [1.0]])
    # solve B^T column by column from  AT @ ((G e_b) * (BT e_a)) = target(a, b)
    BT = np.zeros((n, n))
    for col in range(n):
        rows, rhs = [], []
        for b in range(r):
            gb = G[:, b]
            for i in range(m):
                rows.append(AT[i] * gb)                    # coefficient of BT[:, col]
                rhs.append(1.0 if i + b == col else 0.0)  # y_i = sum_j d[i+j] g[j]
        sol, res, rk, sv = np.linalg.lstsq(np.array(rows), np.array(rhs), rcond=None)
        BT[:, col] = sol
    # check
    rng = np.random.default_rng(1)
    d, g = rng.standard_normal(n), rng.standard_normal(r)
    y = AT @ ((G @ g) * (BT @ d))
    ref = np.array([sum(d[i + j] * g[j] for j in range(r)) for i in range(m)])
    assert np.allclose(y, ref, atol=1e-9), (y, ref)
    return AT, G, BT

if __name__ == '__main__':
    import wino_emul as E
    sets = {'0,1,-1,2,-2': [0, 1, -1, 2, -2], '0,1,-1,1/2,-1/2': [0, 1, -1, Fr(1, 2), Fr(-1, 2)],
            '0,1,-1,2,-1/2': [0, 1, -1, 2, Fr(-1, 2)], '0,1,-1,1/2,-2': [0, 1, -1, Fr(1, 2), -2],
            '0,1,-1,1/2,2': [0, 1, -1, Fr(1, 2), 2], '0,1/2,-1/2,3/2,-3/2': [0, Fr(1, 2), Fr(-1, 2), Fr(3, 2), Fr(-3, 2)]}
    rms = np.sqrt((E.ref ** 2).mean())
    for name, pts in sets.items():
        AT, G, BT = cook_toom(pts, 4, 3)
        F = dict(BT=BT, G=G, AT=AT, m=4)
        y, vmax, mrms = E.wino(E.F23, F, True, act_scale=1.0)
        d = y - E.ref
        print('%-22s err rms %.3e max %.3e (of ref rms)  |V|max %.1f  M rms %.3f  max|BT| %.2f max|AT| %.2f' % (
            name, np.sqrt((d ** 2).mean()) / rms, np.abs(d).max() / rms, vmax, mrms, np.abs(BT).max(), np.abs(AT).max()))
