"""Edge-aligned evaluation of the equivariant kernel on the low-rank path.

The basis of the reference is rotation-covariant, B_f(R r) = D_lo(R) B_f(r) D_li(R)^T (that is what makes ConvSE3
equivariant; B:97-138, 153-205).  With R_e taking the fixed axis a = (0,1,0) -- the polar axis of the reference's spherical
harmonics after its axis permutation (B:57-95) -- to the edge direction,

    out[e,o,:] = D_lo(R_e) * sum_{i,f} R[e,o,i,f] B_f(a) x'[e,i,:],        x'[e,i,:] = D_li(R_e)^T x[j(e),i,:]

and B_f(a) couples only components with |m_p| = |m_q|, as 2x2 blocks (a, -b; b, a): per radial weight 2 FMAs (1 for m = 0)
instead of 2 l_out + 1.  The constant coefficients of B_f(a) are folded into the weight image, so the fused kernel runs
with P = 2 (components +m, -m), F = 2 (weights a_m, b_m) per m >= 1 and P = F = 1 for m = 0.

Host side, float64 torch on the GPU (geometry only, once per forward): rotations, Wigner matrices by evaluating the real
spherical harmonics at rotated sample points (Y(R x_s) = D Y(x_s)), the per-edge "basis" blocks that make se3_tbuild_fwd
produce the rotated neighbour features in the order the kernel consumes them."""
import math

import torch

from . import ops

AXIS = (0.0, 1.0, 0.0)


def real_sh64(d, lmax):
    """Real spherical harmonics of the reference (SH:34-123 with theta = pi - beta, phi = alpha, IR:103-104, and the axis
    permutation (x,y,z) = (c2,c0,c1) of B:57-95), float64, for unit vectors d [N,3] -> list over l of [N, 2l+1], m = -l..l."""
    cx, cy, cz = d[:, 2], d[:, 0], d[:, 1]
    rxy = cx * cx + cy * cy
    beta = torch.atan2(torch.sqrt(rxy), cz)
    phi = torch.atan2(cy, cx)
    ct = torch.cos(math.pi - beta)
    st2 = (1 - ct * ct).clamp(min=0)
    P = {}
    for m in range(lmax + 1):
        if m == 0:
            P[(0, 0)] = torch.ones_like(ct)
        else:
            semif = 1.0
            for k in range(2 * m - 1, 1, -2):
                semif *= k
            P[(m, m)] = ((-1) ** m * semif) * st2.pow(m / 2)
        for l in range(m + 1, lmax + 1):
            y = ((2 * l - 1) / (l - m)) * ct * P[(l - 1, m)]
            if l - m > 1:
                y = y - ((l + m - 1) / (l - m)) * P[(l - 2, m)]
            P[(l, m)] = y
    out = []
    for l in range(lmax + 1):
        comps = []
        for m in range(-l, l + 1):
            ma = abs(m)
            N = math.sqrt((2 * l + 1) / (4 * math.pi))
            if m == 0:
                comps.append(N * P[(l, 0)])
                continue
            poch = 1.0
            for n in range(l - ma + 1, l + ma + 1):
                poch *= n
            N *= math.sqrt(2.0 / poch)
            ang = torch.cos(m * phi) if m > 0 else torch.sin(ma * phi)
            comps.append(ang * P[(l, ma)] * N)
        out.append(torch.stack(comps, dim=-1))
    return out


def rotation_to(rhat):
    """R [E,3,3] float64 with R a = rhat (Rodrigues from a or, for directions in the opposite hemisphere, from -a after a
    half turn about x)."""
    E = rhat.shape[0]
    a = torch.tensor(AXIS, dtype=torch.float64, device=rhat.device)
    flip = (rhat @ a) < 0
    src = torch.where(flip[:, None], -a[None], a[None]).expand(E, 3)
    v = torch.linalg.cross(src, rhat)
    c = (src * rhat).sum(-1)
    vx = torch.zeros((E, 3, 3), dtype=torch.float64, device=rhat.device)
    vx[:, 0, 1], vx[:, 0, 2] = -v[:, 2], v[:, 1]
    vx[:, 1, 0], vx[:, 1, 2] = v[:, 2], -v[:, 0]
    vx[:, 2, 0], vx[:, 2, 1] = -v[:, 1], v[:, 0]
    eye = torch.eye(3, dtype=torch.float64, device=rhat.device)
    R = eye + vx + (vx @ vx) / (1 + c)[:, None, None]
    half_turn = torch.diag(torch.tensor([1.0, -1.0, -1.0], dtype=torch.float64, device=rhat.device))
    return torch.where(flip[:, None, None], R @ half_turn, R)


_SAMPLES = {}


def _samples(l, device):
    key = (l, device)
    if key not in _SAMPLES:
        g = torch.Generator().manual_seed(1000 + l)
        xs = torch.randn(2 * l + 3, 3, generator=g, dtype=torch.float64)
        xs = xs / xs.norm(dim=-1, keepdim=True)
        pin = torch.linalg.pinv(real_sh64(xs, l)[l])          # [M, S]; tiny, on the host (no device SVD launches)
        _SAMPLES[key] = (xs.to(device), pin.to(device))
    return _SAMPLES[key]


def wigner(R, lmax):
    """Representation matrices D_l(R) [E, 2l+1, 2l+1] (float64) in the reference's real basis: Y_l(R x) = D_l(R) Y_l(x)."""
    E = R.shape[0]
    out = [torch.ones((E, 1, 1), dtype=torch.float64, device=R.device)]
    for l in range(1, lmax + 1):
        xs, pin = _samples(l, R.device)
        pts = torch.einsum('eab,sb->esa', R, xs).reshape(-1, 3)
        Yr = real_sh64(pts, l)[l].reshape(E, xs.shape[0], 2 * l + 1)
        out.append(torch.einsum('esm,ns->emn', Yr, pin))
    return out


_COEFFS = {}


def aligned_coeffs(li, lo):
    """B_f(a) of the pair (float64 numpy-free): (c0 [F], ca [min, F], cb [min, F]) with
    c0 = B[lo, li, :], ca[m-1] = B[lo+m, li+m, :] (= B[lo-m, li-m, :]), cb[m-1] = B[lo-m, li+m, :] (= -B[lo+m, li-m, :])."""
    key = (li, lo)
    if key not in _COEFFS:
        a = torch.tensor([AXIS], dtype=torch.float64)
        Y = real_sh64(a, li + lo)
        P, Q, F = 2 * lo + 1, 2 * li + 1, 2 * min(li, lo) + 1
        cols = []
        for J in range(abs(li - lo), li + lo + 1):
            QJ = torch.from_numpy(ops.qj_table(J, li, lo)).double()          # [(P*Q), 2J+1]
            cols.append((Y[J] @ QJ.t()).reshape(P, Q))
        B = torch.stack(cols, dim=-1)                                        # [P, Q, F]
        mn = min(li, lo)
        c0 = B[lo, li, :].clone()
        ca = torch.stack([B[lo + m, li + m, :] for m in range(1, mn + 1)]) if mn else torch.zeros((0, F), dtype=torch.float64)
        cb = torch.stack([B[lo - m, li + m, :] for m in range(1, mn + 1)]) if mn else torch.zeros((0, F), dtype=torch.float64)
        # the structure this formulation relies on
        chk = torch.zeros_like(B)
        chk[lo, li, :] = c0
        for m in range(1, mn + 1):
            chk[lo + m, li + m, :] = ca[m - 1]
            chk[lo - m, li - m, :] = ca[m - 1]
            chk[lo - m, li + m, :] = cb[m - 1]
            chk[lo + m, li - m, :] = -cb[m - 1]
        assert float((chk - B).abs().max()) < 1e-9 * max(1.0, float(B.abs().max())), 'unexpected structure of the aligned basis'
        _COEFFS[key] = (c0, ca, cb)
    return _COEFFS[key]


class EdgeFrames:
    """Per-forward geometry of the aligned formulation: D_l per edge and the tbuild blocks derived from them."""

    def __init__(self, rel_pos, lmax, D=None):
        if D is not None:
            # Wigner matrices already built on the device (ops.frames -> se3_frames_fwd, the production path); this class then
            # only derives the tbuild blocks of the R-first kernels from them
            self.E = rel_pos.numel() // 3
            self.D = [torch.ones((self.E, 1, 1), dtype=torch.float32, device=rel_pos.device)] + list(D[1:])
            self._blocks = {}
            return
        d = rel_pos.reshape(-1, 3).double()
        nrm = d.norm(dim=-1, keepdim=True)
        # coincident points: the reference evaluates its harmonics at beta = atan2(0, 0) = 0, alpha = 0 (B:57-95), which is
        # the direction of the axis a itself, so B(0) = B(a): use the identity frame there
        a = torch.tensor([AXIS], dtype=torch.float64, device=d.device)
        rhat = torch.where(nrm > 0, d / nrm.clamp(min=1e-300), a)
        D = wigner(rotation_to(rhat), lmax)
        self.E = d.shape[0]
        self.D = [m.float().contiguous() for m in D]          # [E, M, M]
        self._blocks = {}

    def block(self, li, m):
        """basis_pair for se3_tbuild_fwd [E, P', Q, F'] so that T'[e,i,f,c] = sum_q block[e,c,q,f] x[j(e),i,q]:
        m = 0: P' = F' = 1: x'[li];   m >= 1: P' = F' = 2: f = a: (x'[+m], x'[-m]),  f = b: (-x'[-m], x'[+m])."""
        key = (li, m)
        if key not in self._blocks:
            D = self.D[li]                                    # x'[n] = sum_q D[q, n] x[q]
            if m == 0:
                blk = D[:, :, li].reshape(self.E, 1, 2 * li + 1, 1)
            else:
                xp, xm = D[:, :, li + m], D[:, :, li - m]    # [E, Q] each
                blk = torch.stack([torch.stack([xp, -xm], dim=-1),       # c = 0 (component +m): f = a -> x'[+m], f = b -> -x'[-m]
                                   torch.stack([xm, xp], dim=-1)], dim=1)  # c = 1 (component -m): f = a -> x'[-m], f = b -> x'[+m]
            self._blocks[key] = blk.contiguous().reshape(-1)
        return self._blocks[key]
