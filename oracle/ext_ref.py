"""Test infrastructure only: numpy restatement of the OPTIONAL physics of st_step_ext
(sitrack_b200/csrc/st_ext.cu).  The reference has no such modes (si3_part_tracker.py:423-484 is Euler + face
pick + one hop), so this is not a parity oracle: it restates the same definitions independently --
orientation walk over convex cells, C-grid linear interpolation between the face points, Euler / midpoint /
classical Runge-Kutta on the frozen record, Survive on every cell entered -- so that the CUDA kernel can be
checked on curvilinear grids where no closed form exists.  Plain float64 numpy, one buoy at a time.
"""
import numpy as np


def _orient(ay, ax, by, bx, cy, cx):
    return (bx - ax) * (cy - ay) - (by - ay) * (cx - ax)


def _moves(g, y, x, j, i):
    """(dj, di) of the orientation walk for point (y,x) seen from cell (j,i)."""
    Yf, Xf = g["Yf"], g["Xf"]
    bl = (Yf[j - 1, i - 1], Xf[j - 1, i - 1]); br = (Yf[j - 1, i], Xf[j - 1, i])
    ur = (Yf[j, i], Xf[j, i]); ul = (Yf[j, i - 1], Xf[j, i - 1])
    dj = int(_orient(*ur, *ul, y, x) < 0.0) - int(_orient(*bl, *br, y, x) < 0.0)
    di = int(_orient(*br, *ur, y, x) < 0.0) - int(_orient(*ul, *bl, y, x) < 0.0)
    return dj, di


def walk_to(g, y, x, j, i, max_hops):
    Nj, Ni = g["tmask"].shape
    h = 0
    while True:
        dj, di = _moves(g, y, x, j, i)
        if dj == 0 and di == 0:
            return j, i
        if h >= max_hops:
            return j, i
        jn, in_ = min(max(j + dj, 1), Nj - 2), min(max(i + di, 1), Ni - 2)
        if (jn, in_) == (j, i):
            return j, i
        j, i = jn, in_
        h += 1


def _ccw(A, B, C):
    return (C[0] - A[0]) * (B[1] - A[1]) > (B[0] - A[0]) * (C[1] - A[1])


def _intersect(A, B, C, D):
    return (_ccw(A, C, D) != _ccw(B, C, D)) and (_ccw(A, B, C) != _ccw(A, B, D))


def velocity_at(g, U, V, y, x, j, i, interp, uv_strategy=1):
    uW, uE = float(U[j, i - 1]), float(U[j, i])
    vS, vN = float(V[j - 1, i]), float(V[j, i])
    if interp == 1:
        pw = (g["Yu"][j, i - 1], g["Xu"][j, i - 1]); pe = (g["Yu"][j, i], g["Xu"][j, i])
        ps = (g["Yv"][j - 1, i], g["Xv"][j - 1, i]); pn = (g["Yv"][j, i], g["Xv"][j, i])
        ex, ey = pe[1] - pw[1], pe[0] - pw[0]
        nx, ny = pn[1] - ps[1], pn[0] - ps[0]
        xi = ((x - pw[1]) * ex + (y - pw[0]) * ey) / (ex * ex + ey * ey)
        et = ((x - ps[1]) * nx + (y - ps[0]) * ny) / (nx * nx + ny * ny)
        xi, et = min(max(xi, 0.0), 1.0), min(max(et, 0.0), 1.0)
        return uW + xi * (uE - uW), vS + et * (vN - vS)
    if uv_strategy == 1:
        P = (y, x); Fp = (g["Yf"][j, i], g["Xf"][j, i])
        lu = _intersect(P, Fp, (g["Yv"][j - 1, i], g["Xv"][j - 1, i]), (g["Yv"][j, i], g["Xv"][j, i]))
        lv = _intersect(P, Fp, (g["Yu"][j, i - 1], g["Xu"][j, i - 1]), (g["Yu"][j, i], g["Xu"][j, i]))
        return (uW if lu else uE), (vS if lv else vN)
    return 0.5 * (uE + uW), 0.5 * (vN + vS)


def killed(g, IC, j, i, rmin_conc=0.1):
    tm = g["tmask"]
    Nj, Ni = tm.shape
    if j <= 1 or j >= Nj - 2 or i <= 1 or i >= Ni - 2:
        return True
    if int(tm[j, i]) + int(tm[j, i + 1]) + int(tm[j + 1, i]) + int(tm[j, i - 1]) + int(tm[j - 1, i - 1]) < 5:
        return True
    s = float(IC[j, i]) + float(IC[j, i + 1]); s += float(IC[j + 1, i]); s += float(IC[j, i - 1]); s += float(IC[j - 1, i - 1])
    return 0.2 * s < rmin_conc


def step(g, U, V, IC, pos, cell, alive, scheme, interp, max_hops, rdt=3600.0, uv_strategy=1):
    """One record for every buoy, in place; returns the row (yx, mask)."""
    h = rdt / 1000.0
    nP = pos.shape[0]
    yx = np.full((nP, 2), -9999.0); mk = np.zeros(nP, np.int8)
    for b in range(nP):
        if not alive[b]:
            continue
        y, x = pos[b]; j, i = int(cell[b, 0]), int(cell[b, 1])
        k1u, k1v = velocity_at(g, U, V, y, x, j, i, interp, uv_strategy)
        du, dv = k1u, k1v
        if scheme == 2:
            qy, qx = y + 0.5 * h * k1v, x + 0.5 * h * k1u
            j2, i2 = walk_to(g, qy, qx, j, i, max_hops)
            du, dv = velocity_at(g, U, V, qy, qx, j2, i2, interp, uv_strategy)
        elif scheme == 4:
            qy, qx = y + 0.5 * h * k1v, x + 0.5 * h * k1u
            js, is_ = walk_to(g, qy, qx, j, i, max_hops)
            k2u, k2v = velocity_at(g, U, V, qy, qx, js, is_, interp, uv_strategy)
            qy, qx = y + 0.5 * h * k2v, x + 0.5 * h * k2u
            js, is_ = walk_to(g, qy, qx, js, is_, max_hops)
            k3u, k3v = velocity_at(g, U, V, qy, qx, js, is_, interp, uv_strategy)
            qy, qx = y + h * k3v, x + h * k3u
            js, is_ = walk_to(g, qy, qx, js, is_, max_hops)
            k4u, k4v = velocity_at(g, U, V, qy, qx, js, is_, interp, uv_strategy)
            du = (k1u + 2.0 * k2u + 2.0 * k3u + k4u) * (1.0 / 6.0)
            dv = (k1v + 2.0 * k2v + 2.0 * k3v + k4v) * (1.0 / 6.0)
        yn, xn = y + h * dv, x + h * du
        yx[b] = (yn, xn); mk[b] = 1
        for _ in range(max_hops):
            dj, di = _moves(g, yn, xn, j, i)
            if dj == 0 and di == 0:
                break
            j += dj; i += di
            if killed(g, IC, j, i):
                alive[b] = 0
                break
        pos[b] = (yn, xn); cell[b] = (j, i)
    return yx, mk
