"""Drop-in for `sitrack.tracking` (reference: sitrack/tracking.py): same names,
argument order, return shapes and in-place behaviour.  Geometry arithmetic runs on
the GPU (libsitrack_b200.so); index bookkeeping and array shrinking stay in numpy.
"""
import numpy as np

from . import _lib, config

__all__ = ['rmin_conc', 'rFoundKM', 'GetTimeSpan', 'intersect2Seg', 'intersect2SegBatch', 'Survive', 'SurviveBatch', 'SeedInit', 'CrossedEdge', 'NewHostCell', 'UpdtInd4NewCell', 'debugSeeding', 'SidfexSeeding', 'ReadFromSidfexDatFile', 'nemoSeed']
from ._lib import as_c, check, hptr

rmin_conc = 0.1    # tracking.py:4 ice concentration below which a buoy is discontinued
rFoundKM = 2.5     # tracking.py:5


def GetTimeSpan(dt, vtime_mod, iSdA, iMdA, iMdB, iStop=None, iverbose=0):
    """tracking.py:8-37 -> Nt, kt0, ktN, itM0, itMN (host logic on the time axis)."""
    from .util import epoch2clock as e2c
    if iSdA < iMdA - dt / 2 or iSdA > iMdB - dt / 2:
        print('PROBLEM: time in the seeding file (' + e2c(iSdA) + ') is outside of what model spans!')
        raise SystemExit(0)
    vt = np.asarray(vtime_mod)
    kt0 = int(np.argmin(np.abs(vt - iSdA)))
    if iSdA >= vt[kt0]:
        kt0 += 1
    ktN = int(np.argmin(np.abs(vt - iStop))) if iStop else len(vt) - 1
    itM0, itMN = vt[kt0], vt[ktN]
    Nt = ktN - kt0 + 1
    print('    * [GetTimeSpan]: First record needed =', kt0, 'of SI3 file =>', e2c(itM0))
    print('    * [GetTimeSpan]: Last record needed =', ktN, 'of SI3 file =>', e2c(itMN))
    print('       ==> ' + str(Nt) + ' model records')
    print('       ==> that makes ' + str(round((itMN - itM0) / (3600 * 24), 3)) + ' days of ice particule tracking.')
    return Nt, kt0, ktN, itM0, itMN


def intersect2SegBatch(A, B, C, D):
    """(n,2) x4 [y,x] -> bool (n,): do segments AB and CD intersect (tracking.py:51-58)."""
    a, b, c, d = (as_c(p, np.float64).reshape(-1, 2) for p in (A, B, C, D))
    out = np.zeros(a.shape[0], np.int8)
    check(_lib.lib().st_intersect2seg(config.device, a.shape[0], hptr(a), hptr(b), hptr(c), hptr(d), hptr(out)))
    return out.astype(bool)


def intersect2Seg(pcA, pcB, pcC, pcD):
    return bool(intersect2SegBatch([pcA], [pcB], [pcC], [pcD])[0])


def _stencil5(a, jT, iT):
    # [j,i] [j,i+1] [j+1,i] [j,i-1] [j-1,i-1]  (tracking.py:79,87)
    return [a[jT, iT], a[jT, iT + 1], a[jT + 1, iT], a[jT, iT - 1], a[jT - 1, iT - 1]]


def SurviveBatch(kjiT, pmskT, pIceC):
    """kjiT (n,2) -> kill flags (n,) int; pIceC a (Nj,Ni) array."""
    Nj, Ni = np.shape(pmskT)
    ji = as_c(kjiT, np.int32).reshape(-1, 2)
    n = ji.shape[0]
    jj = np.clip(ji[:, 0], 1, Nj - 2); ii = np.clip(ji[:, 1], 1, Ni - 2)     # border cells die in test 1 anyway
    tm5 = as_c(np.stack(_stencil5(np.asarray(pmskT), jj, ii), axis=1), np.int8)
    ic5 = as_c(np.stack(_stencil5(np.asarray(pIceC), jj, ii), axis=1), np.float64)
    out = np.zeros(n, np.int32)
    check(_lib.lib().st_survive(config.device, n, hptr(ji), Nj, Ni, hptr(tm5), hptr(ic5), rmin_conc, hptr(out)))
    return out


def Survive(kID, kjiT, pmskT, pIceC=[], iverbose=0):
    """tracking.py:62-93 -> 0 keep / 1 discontinue (domain edge, land, low concentration)."""
    if len(np.shape(pIceC)) != 2:
        # the reference reaches `if zic < rmin_conc` with zic never assigned (tracking.py:86-89)
        raise UnboundLocalError("Survive: a 2-D `pIceC` is required (the reference fails the same way)")
    ikill = int(SurviveBatch([kjiT], pmskT, pIceC)[0])
    if ikill and iverbose > 0:
        print('        ===> I CANCEL buoy ' + str(kID) + '!!!')
    return ikill


def SeedInit(pIDs, pSG, pSC, platT, plonT, pYf, pXf, pResolKM, maskT, xIceConc=[], iverbose=0):
    """tracking.py:98-178 -> nP, pSG', pSC', IDs', zjiT (nP,2), zJIvrt (nP,2,4), iKeep."""
    from time import time
    from .engine import TrackEngine
    (nP, n2) = np.shape(pSG)
    if np.shape(pSC) != (nP, n2):
        print('ERROR [SeedInit]: shape disagreement for `pSG` and `pSC`!'); raise SystemExit(0)
    if n2 != 2:
        print('ERROR [SeedInit]: wrong shape for `pSG` and `pSC`!'); raise SystemExit(0)
    if len(np.shape(xIceConc)) != 2:
        raise UnboundLocalError("SeedInit: a 2-D `xIceConc` is required (the reference's Survive fails without it)")
    t0 = time()
    with TrackEngine(pYf, pXf, tmask=maskT, uv_strategy=0, rmin_conc=rmin_conc, device=config.device) as eng:
        eng.set_locate_grid(platT, plonT, pResolKM)
        cell, near, kmask = eng.seed_locate(pSG, pSC, np.asarray(xIceConc, np.float32))
    print(' * [SeedInit]: number of seconds it took to locate all the points in a target grid cell:', time() - t0)
    zjiT = cell.astype(int)
    zJIvrt = np.zeros((nP, 2, 4), dtype=int)
    zJIvrt[:, 0, :] = zjiT[:, 0:1] + np.array([-1, -1, 0, 0])
    zJIvrt[:, 1, :] = zjiT[:, 1:2] + np.array([-1, 0, 0, -1])
    zjiT[kmask == 0] = 0; zJIvrt[kmask == 0] = 0
    iKeep = np.arange(nP, dtype=int)
    nPn = int(np.sum(kmask))
    if nPn < nP:
        (iGone,) = np.where(kmask == 0)
        print(' * [SeedInit()]: ' + str(nP - nPn) + ' "to-be-seeded" buoys have to be canceled.')
        print('          => their IDs:', np.asarray(pIDs)[iGone])
        nP = nPn
        (iKeep,) = np.where(kmask == 1)
    return nP, pSG[iKeep, :], pSC[iKeep, :], pIDs[iKeep], zjiT[iKeep, :], zJIvrt[iKeep, :, :], iKeep


def _ring(ji4vert, pY, pX):
    """12 F-points around a cell: its 4 vertices then the 8 outward neighbours NewHostCell looks at."""
    [[jbl, jbr, jur, jul], [ibl, ibr, iur, iul]] = np.asarray(ji4vert)
    idx = [(jbl, ibl), (jbr, ibr), (jur, iur), (jul, iul),
           (jbl - 1, ibl), (jbr - 1, ibr), (jbr, ibr + 1), (jur, iur + 1),
           (jul + 1, iul), (jur + 1, iur), (jul, iul - 1), (jbl, ibl - 1)]
    return np.array([[pY[j, i], pX[j, i]] for (j, i) in idx], np.float64)


def _walk(pP1, pP2, ji4vert, pY, pX, kcross_in=None):
    p1 = as_c([pP1], np.float64); p2 = as_c([pP2], np.float64)
    ring = as_c(_ring(ji4vert, pY, pX)[None], np.float64)
    kin = None if kcross_in is None else as_c([kcross_in], np.int32)
    kc = np.zeros(1, np.int32); kn = np.zeros(1, np.int32)
    check(_lib.lib().st_cell_walk(config.device, 1, hptr(p1), hptr(p2), hptr(ring), hptr(kin), hptr(kc), hptr(kn)))
    return int(kc[0]), int(kn[0])


def CrossedEdge(pP1, pP2, ji4vert, pY, pX, iverbose=0):
    """tracking.py:182-200 -> 1 bottom, 2 right, 3 top, 4 left (4 also when no edge is met)."""
    return _walk(pP1, pP2, ji4vert, pY, pX)[0]


def NewHostCell(kcross, pP1, pP2, ji4vert, pY, pX, iverbose=0):
    """tracking.py:203-249 -> 1..4 edge neighbours, 5..8 diagonal neighbours."""
    if kcross not in (1, 2, 3, 4):
        return kcross
    return _walk(pP1, pP2, ji4vert, pY, pX, kcross_in=kcross)[1]


_SHIFT = {1: (-1, 0), 2: (0, 1), 3: (1, 0), 4: (0, -1), 5: (-1, -1), 6: (-1, 1), 7: (1, 1), 8: (1, -1)}


def UpdtInd4NewCell(knhc, ji4vert, kjiT, iverbose=0):
    """tracking.py:253-305 -- shifts the vertex and T indices IN PLACE and returns them."""
    if knhc not in _SHIFT:
        print('ERROR: unknown direction, knhc=', knhc)
        raise SystemExit(0)
    dj, di = _SHIFT[knhc]
    if dj:
        ji4vert[0, :] = ji4vert[0, :] + dj
        kjiT[0] = kjiT[0] + dj
    if di:
        ji4vert[1, :] = ji4vert[1, :] + di
        kjiT[1] = kjiT[1] + di
    return ji4vert, kjiT


# ---- seeding helpers (tracking.py:313-442): plain array masking, host side -----------------

def debugSeeding():
    return np.array([[84., 20.], [89., 50.], [63., -11.], [85., 100.], [89., 100.], [79., 180.], [76., 46.],
                     [75., 190.], [85.2, -15.], [75., 210.], [75., -72.], [83., 200.], [79., -42.], [85., 300.]])


def ReadFromSidfexDatFile(filepath='./sidfexloc.dat'):
    from os.path import exists
    if not exists(filepath):
        print("=== Error ! 'sidfexloc.dat' file is missing.")
        raise SystemExit(0)
    return np.genfromtxt(open(filepath))


def SidfexSeeding(filepath='./sidfexloc.dat'):
    dat = ReadFromSidfexDatFile(filepath)
    return dat[:, [2, 1]], dat[:, 0].astype(int)


def nemoSeed(pmskT, platT, plonT, pIC, khss=1, fmsk_rstrct=[], platF=[], plonF=[]):
    """tracking.py:365-442: every khss-th ocean T-point north of 55N over ice >= 0.9
    (optionally F-points whose 4 T neighbours qualify)."""
    sl = (slice(None, None, khss), slice(None, None, khss))
    msk = np.array(pmskT[sl], dtype='i1')
    if np.shape(fmsk_rstrct) == np.shape(pmskT):
        msk = msk * np.asarray(fmsk_rstrct)[sl]
    lat, lon = np.asarray(platT)[sl], np.asarray(plonT)[sl]
    msk[lat < 55.] = 0
    msk[np.asarray(pIC, dtype=float)[sl] < 0.9] = 0
    jj, ii = np.where(msk == 1)
    out = np.stack([lat[jj, ii], lon[jj, ii]], axis=1).astype(float)
    if np.shape(platF) == np.shape(pmskT) and np.shape(plonF) == np.shape(pmskT):
        mF = np.zeros(msk.shape, dtype='i1')
        mF[1:-1, 1:-1] = (msk[2:, 1:-1] + msk[1:-1, 2:] + msk[:-2, 1:-1] + msk[1:-1, :-2]) / 4
        jf, jf_i = np.where(mF == 1)
        outF = np.stack([np.asarray(platF)[sl][jf, jf_i], np.asarray(plonF)[sl][jf, jf_i]], axis=1).astype(float)
        print(' * [nemoSeed()]: adding ', len(jf), 'F-points to the', len(jj), 'T-points!')
        out = np.concatenate([out, outF])
    return out
