"""Pure-Python port of the reference's record x buoy loop and its helpers.

TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's cpu_baseline / --impl reference).
The upstream tracker is interpreted Python over numpy scalars; this port keeps that
execution model (one interpreted iteration per buoy per record, numpy 2-D indexing for
every grid access) so that timing it on the GPU box's host cores states what a user of
the reference gets there -- /root/reference itself cannot travel to that box.  Pinned
bit-exact against the reference's own functions by tests/golden/ (see make_golden.py).
Reference lines: si3_part_tracker.py:361-496, sitrack/tracking.py:44-305,
sitrack/locate.py:49-78.
"""
import numpy as np

FILL = -9999.0
RMIN_CONC = 0.1


def ccw(A, B, C):                                   # tracking.py:44-49
    return (C[0] - A[0]) * (B[1] - A[1]) > (B[0] - A[0]) * (C[1] - A[1])


def intersect(A, B, C, D):                          # tracking.py:51-58
    return (ccw(A, C, D) != ccw(B, C, D)) and (ccw(A, B, C) != ccw(A, B, D))


def inside_quad(y, x, quad):                        # locate.py:49-78
    inside = False
    xints = 0.0
    y1, x1 = quad[0]
    for i in range(5):
        y2, x2 = quad[i % 4]
        if y > min(y1, y2):
            if y <= max(y1, y2):
                if x <= max(x1, x2):
                    if y1 != y2:
                        xints = (y - y1) * (x2 - x1) / (y2 - y1) + x1
                    if x1 == x2 or x <= xints:
                        inside = not inside
        y1, x1 = y2, x2
    return inside


def survive(jT, iT, tmask, ic):                     # tracking.py:62-93
    Nj, Ni = tmask.shape
    if jT in (0, 1, Nj - 2, Nj - 1) or iT in (0, 1, Ni - 2, Ni - 1):
        return 1
    if tmask[jT, iT] + tmask[jT, iT + 1] + tmask[jT + 1, iT] + tmask[jT, iT - 1] + tmask[jT - 1, iT - 1] < 5:
        return 1
    zic = 0.2 * (ic[jT, iT] + ic[jT, iT + 1] + ic[jT + 1, iT] + ic[jT, iT - 1] + ic[jT - 1, iT - 1])
    return 1 if zic < RMIN_CONC else 0


def _pt(Y, X, j, i):
    return [Y[j, i], X[j, i]]


def crossed_edge(P1, P2, jv, iv, Y, X):             # tracking.py:182-200
    for kk in range(4):
        k1 = (kk + 1) % 4
        if intersect(P1, P2, _pt(Y, X, jv[kk], iv[kk]), _pt(Y, X, jv[k1], iv[k1])):
            return kk + 1
    return 4


_OUTWARD = {  # crossed edge -> ((vertex, dj, di, answer), (vertex, dj, di, answer)); tracking.py:215-243
    1: ((0, -1, 0, 5), (1, -1, 0, 6)),
    2: ((1, 0, 1, 6), (2, 0, 1, 7)),
    3: ((3, 1, 0, 8), (2, 1, 0, 7)),
    4: ((3, 0, -1, 8), (0, 0, -1, 5)),
}
_SHIFT = {1: (-1, 0), 2: (0, 1), 3: (1, 0), 4: (0, -1), 5: (-1, -1), 6: (-1, 1), 7: (1, 1), 8: (1, -1)}


def new_host_cell(kcross, P1, P2, jv, iv, Y, X):    # tracking.py:203-249
    for (v, dj, di, ans) in _OUTWARD[kcross]:
        if intersect(P1, P2, _pt(Y, X, jv[v], iv[v]), _pt(Y, X, jv[v] + dj, iv[v] + di)):
            return ans
    return kcross


def advance(g, xU, xV, xIC, jrec, cur, nxt, mnxt, jiT, alive, first, last, vMesh, still_in,
            uv_strategy=1, rdt=3600.0):
    """One record for every buoy: the body of si3_part_tracker.py:378-488.  xU/xV/xIC are the
    f8 work arrays of the record; cur/nxt rows jt and jt+1 of xPosC.  Returns buoy-steps done."""
    Yf, Xf, Yu, Xu, Yv, Xv, tmask = (g[k] for k in ("Yf", "Xf", "Yu", "Xu", "Yv", "Xv", "tmask"))
    nsteps = 0
    for b in range(cur.shape[0]):
        if alive[b] == 1 and jrec >= first[b] and jrec <= last[b]:
            ry, rx = cur[b, :]
            jT, iT = jiT[b, :]
            if uv_strategy == 0:
                zU = 0.5 * (xU[jT, iT] + xU[jT, iT - 1])
                zV = 0.5 * (xV[jT, iT] + xV[jT - 1, iT])
            else:
                Fp = _pt(Yf, Xf, jT, iT)
                zU = xU[jT, iT - 1] if intersect([ry, rx], Fp, _pt(Yv, Xv, jT - 1, iT), _pt(Yv, Xv, jT, iT)) else xU[jT, iT]
                zV = xV[jT - 1, iT] if intersect([ry, rx], Fp, _pt(Yu, Xu, jT, iT - 1), _pt(Yu, Xu, jT, iT)) else xV[jT, iT]
            dx = zU * rdt
            dy = zV * rdt
            rxn = rx + dx / 1000.
            ryn = ry + dy / 1000.
            nxt[b, :] = [ryn, rxn]
            mnxt[b] = 1
            nsteps += 1
            jv = [jT - 1, jT - 1, jT, jT]; iv = [iT - 1, iT, iT, iT - 1]
            if not still_in[b]:                    # cached cell corners, reloaded after a cell change (:391-402)
                vMesh[b, :, :] = [_pt(Yf, Xf, jv[k], iv[k]) for k in range(4)]
            still_in[b] = inside_quad(ryn, rxn, vMesh[b, :, :])
            if not still_in[b]:
                kc = crossed_edge([ry, rx], [ryn, rxn], jv, iv, Yf, Xf)
                kn = new_host_cell(kc, [ry, rx], [ryn, rxn], jv, iv, Yf, Xf)
                jiT[b, 0] += _SHIFT[kn][0]; jiT[b, 1] += _SHIFT[kn][1]
                if survive(jiT[b, 0], jiT[b, 1], tmask, xIC) > 0:
                    alive[b] = 0
    return nsteps


def track(g, U, V, IC, pos0, jiT0, kstrt=0, rec_first=None, rec_last=None, uv_strategy=1, rdt=3600.0):
    """-> posC (nrec+1,nP,2), mask (nrec+1,nP), jiT (nP,2), alive (nP,), n buoy-steps done."""
    tmask = g["tmask"]
    nrec = U.shape[0]
    nP = pos0.shape[0]
    first = np.zeros(nP, int) + kstrt if rec_first is None else np.asarray(rec_first, int)
    last = np.zeros(nP, int) + kstrt + nrec - 1 if rec_last is None else np.asarray(rec_last, int)
    posC = np.zeros((nrec + 1, nP, 2)) + FILL
    mask = np.zeros((nrec + 1, nP), 'i1')
    for b in range(nP):
        posC[first[b] - kstrt, b, :] = pos0[b, :]
        mask[first[b] - kstrt, b] = 1
    jiT = np.array(jiT0, dtype=int).copy()
    alive = np.zeros(nP, 'i1') + 1
    (Nj, Ni) = tmask.shape
    xU, xV, xIC = np.zeros((Nj, Ni)), np.zeros((Nj, Ni)), np.zeros((Nj, Ni))
    vMesh = np.zeros((nP, 4, 2))
    still_in = np.zeros(nP, dtype=bool)
    nsteps = 0
    for jt in range(nrec):
        xIC[:, :] = IC[jt]; xU[:, :] = U[jt]; xV[:, :] = V[jt]          # f4 -> f8 (:372-374)
        nsteps += advance(g, xU, xV, xIC, jt + kstrt, posC[jt], posC[jt + 1], mask[jt + 1], jiT, alive,
                          first, last, vMesh, still_in, uv_strategy, rdt)
    return posC, mask, jiT, alive, nsteps
