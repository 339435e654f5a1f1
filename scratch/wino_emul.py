"""Scratch: numerical emulation of Winograd variants for the 3x3 gate conv with (hi, lo) fp16 operand pairs and a
truncating fp32 accumulator (the tensor core), against float64.  F(2x2), F(2x4), F(4x4)."""
import numpy as np, sys
rng = np.random.default_rng(0)
C, CO, H, W = 512, 48, 30, 40
h = (rng.uniform(-1, 1, (C, H, W)) * rng.uniform(0, 1, (C, 1, 1))).astype(np.float32)
# 22-bit representable like the cell kernel's split
def split(x, scale):
    xs = x.astype(np.float64) * scale
    hi = xs.astype(np.float16).astype(np.float64)
    lo = (xs - hi).astype(np.float16).astype(np.float64)
    return hi, lo
hh, hl = split(h, 1.0); h = ((hh + hl / 1.0)).astype(np.float64)   # h exactly hi+lo
w = (rng.standard_normal((CO, C, 3, 3)) * 0.02).astype(np.float32).astype(np.float64)
hp = np.pad(h, ((0, 0), (1, 1), (1, 1)))
ref = np.zeros((CO, H, W))
for ky in range(3):
    for kx in range(3):
        ref += np.einsum('oc,chw->ohw', w[:, :, ky, kx], hp[:, ky:ky + H, kx:kx + W])

BIAS_FIX = 5.5e-7


def trunc32(x):
    """fp64 -> fp32 truncating toward zero"""
    y = x.astype(np.float32)
    bad = np.abs(y.astype(np.float64)) > np.abs(x)
    y[bad] = np.nextafter(y[bad], np.float32(0))
    return y

def tc_gemm(Uh, Ul, Wh, Wl, two_acc=False):
    """[R,K] x [CO,K] -> [R,CO]: per k16 step acc = trunc32(acc + exact partial)"""
    R, K = Uh.shape
    acc = np.zeros((R, Wh.shape[0]), np.float32)
    acc2 = np.zeros_like(acc)
    for k in range(0, K, 16):
        s = slice(k, k + 16)
        hl_ = Uh[:, s] @ Wl[:, s].T; lh = Ul[:, s] @ Wh[:, s].T; hh_ = Uh[:, s] @ Wh[:, s].T
        if two_acc:
            acc2 = trunc32(acc2.astype(np.float64) + hl_); acc2 = trunc32(acc2.astype(np.float64) + lh)
            acc = trunc32(acc.astype(np.float64) + hh_)
        else:
            acc = trunc32(acc.astype(np.float64) + hl_); acc = trunc32(acc.astype(np.float64) + lh)
            acc = trunc32(acc.astype(np.float64) + hh_)
    # the drain warps multiply the truncation bias of the main accumulator back (decoder.cuh kAccTruncFix)
    return (acc.astype(np.float64) * (1 + BIAS_FIX) + acc2.astype(np.float64)) if two_acc else acc.astype(np.float64)

F23 = dict(BT=np.array([[1, 0, -1, 0], [0, 1, 1, 0], [0, -1, 1, 0], [0, 1, 0, -1]], float),
           G=np.array([[1, 0, 0], [.5, .5, .5], [.5, -.5, .5], [0, 0, 1]], float),
           AT=np.array([[1, 1, 1, 0], [0, 1, -1, -1]], float), m=2)
F43 = dict(BT=np.array([[4, 0, -5, 0, 1, 0], [0, -4, -4, 1, 1, 0], [0, 4, -4, -1, 1, 0], [0, -2, -1, 2, 1, 0],
                        [0, 2, -1, -2, 1, 0], [0, 4, 0, -5, 0, 1]], float),
           G=np.array([[1 / 4, 0, 0], [-1 / 6, -1 / 6, -1 / 6], [-1 / 6, 1 / 6, -1 / 6], [1 / 24, 1 / 12, 1 / 6],
                       [1 / 24, -1 / 12, 1 / 6], [0, 0, 1]], float),
           AT=np.array([[1, 1, 1, 1, 1, 0], [0, 1, -1, 2, -2, 0], [0, 1, 1, 4, 4, 0], [0, 1, -1, 8, -8, 1]], float), m=4)

def wino(Fy, Fx, two_acc=False, act_scale=256.0):
    my, mx = Fy['m'], Fx['m']; ay, ax = my + 2, mx + 2
    ty, tx = H // my, W // mx
    Hp = (H + my - 1) // my * my; ty = Hp // my
    assert W % mx == 0
    # weights: G g G^T in fp64, split with per-tensor scale
    U_w = np.einsum('ia,ocab,jb->ijoc', Fy['G'], w, Fx['G'])            # [ay,ax,CO,C]
    sw = 2.0 ** np.floor(np.log2(30000.0 / np.abs(U_w).max()))
    # input transform in fp32
    hp32 = np.pad(h.astype(np.float32), ((0, 0), (1, 1 + (H + my - 1) // my * my - H), (1, 1)))
    tiles = np.zeros((ty, tx, ay, ax, C), np.float32)
    for a in range(ay):
        for b in range(ax):
            tiles[:, :, a, b, :] = hp32[:, a:a + my * ty:my, b:b + mx * tx:mx].transpose(1, 2, 0)
    BTy, BTx = Fy['BT'].astype(np.float32), Fx['BT'].astype(np.float32)
    t1 = np.einsum('ia,yxabc->yxibc', BTy, tiles).astype(np.float32)   # note: einsum accumulates in fp32 here
    V = np.einsum('jb,yxibc->yxijc', BTx, t1).astype(np.float32)
    out = np.zeros((CO, H, W))
    M = np.zeros((ay, ax, ty * tx, CO))
    for i in range(ay):
        for j in range(ax):
            Uh, Ul = split(V[:, :, i, j, :].reshape(ty * tx, C), act_scale)
            Wh, Wl = split(U_w[i, j], sw)
            M[i, j] = tc_gemm(Uh, Ul, Wh, Wl, two_acc) / (sw * act_scale)
    M32 = M.astype(np.float32)
    t = np.einsum('ri,ijtc->rjtc', Fy['AT'].astype(np.float32), M32).astype(np.float32)
    Y = np.einsum('sj,rjtc->rstc', Fx['AT'].astype(np.float32), t).astype(np.float32)   # [my,mx,T,CO]
    Y = Y.reshape(my, mx, ty, tx, CO).transpose(4, 2, 0, 3, 1).reshape(CO, Hp, W)[:, :H]
    return Y.astype(np.float64), float(np.abs(V).max()), float(np.sqrt((M ** 2).mean()))

rms = np.sqrt((ref ** 2).mean())
if __name__ != "__main__":
    pass
if __name__ == '__main__':
    print('ref rms %.4f max %.3f' % (rms, np.abs(ref).max()))
    for name, Fy, Fx, ta in [('F(2x2) 1acc', F23, F23, False), ('F(2x2) 2acc', F23, F23, True), ('F(2x4) 1acc', F23, F43, False),
                             ('F(2x4) 2acc', F23, F43, True), ('F(4x4) 2acc', F43, F43, True), ('F(4x4) 1acc', F43, F43, False)]:
        y, vmax, mrms = wino(Fy, Fx, ta, act_scale=1.0)
        d = y - ref
        print('%-14s err rms %.3e  max %.3e  (rel to ref rms: %.3e / %.3e)  mean signed err*sign(ref) %.3e  |V|max %.1f  M rms %.3f' % (
            name, np.sqrt((d ** 2).mean()), np.abs(d).max(), np.sqrt((d ** 2).mean()) / rms, np.abs(d).max() / rms,
            (d * np.sign(ref)).mean() / rms, vmax, mrms))
