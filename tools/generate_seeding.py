#!/usr/bin/env python3
"""Seeding-file generator: the `tools/generate_idealized_seeding.py` / `generate_sidfex_seeding.py`
workflow of the reference on top of sitrack_b200 (no mojito / cartopy / gudhi).

    python tools/generate_seeding.py -d 1996-12-15_00:00:00 -m mesh_mask.nc -i icemod.nc -k 0 -S 5
    python tools/generate_seeding.py -d 1996-12-15_00:00:00 -m mesh_mask.nc --lsidfex 1 --sidfexdat sidfexloc.dat

Same flags as upstream (-d -m -i -v -k -S -f -C -N, --lsidfex) and the same output name
`./nc/sitrack_seeding_<type>_<YYYYMMDD_hh>[_HSSn].nc` that the tracker parses
(reference si3_part_tracker.py:117-127, ncio.py:333-335).  Inputs may be netCDF or the `.npz`
equivalents; with `.npz` inputs the output is `.npz` too.  The geographic -> km conversion of the
seeds runs on the GPU (Geo2CartNPSkm1D).  Not provided: -C coarsening (gudhi sub-sampling) and the
coastal cleaning that needs the dist2coast data set.
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sitrack_b200 as sit          # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser(description='SITRACK seeding generator (sitrack_b200)')
    ap.add_argument('-d', '--dat0', required=True, help='initial date in the form <YYYY-MM-DD_hh:mm:ss>')
    ap.add_argument('-m', '--fmmm', default=None, help='model `mesh_mask` file of NEMO config used in SI3 run')
    ap.add_argument('-i', '--fsi3', default=None, help='output file of SI3 containing sea-ice concentration')
    ap.add_argument('-v', '--nsic', default='siconc', help='name of sea-ice concentration in SI3 file')
    ap.add_argument('-k', '--krec', type=int, default=0, help='use sea-ice concentration at this record')
    ap.add_argument('-S', '--ihss', type=int, default=1, help='horizontal subsampling factor to apply')
    ap.add_argument('-f', '--fmsk', default=None, help='mask (on SI3 model domain) to control seeding region')
    ap.add_argument('-C', '--crsn', type=int, default=0, help='(not provided) coarsening in km')
    ap.add_argument('-N', '--ncnf', default='NANUK4', help='name of the horizontal NEMO config used')
    ap.add_argument('--lsidfex', type=int, default=0, help='1: SIDFEX seeding from a text file of `id lon lat` rows')
    ap.add_argument('--sidfexdat', default='./sidfexloc.dat', help='SIDFEX text file (with --lsidfex 1)')
    a = ap.parse_args(argv)

    if a.fsi3 and not a.fmmm:
        print('ERROR: you have to specify a MeshMask file with `-m` when using SI3 file!')
        raise SystemExit(0)
    if a.crsn >= 1:
        raise NotImplementedError("-C coarsening relies on gudhi's sparsify_point_set and is not part of sitrack_b200")
    if a.ihss < 1 or a.ihss > 20:
        print('ERROR: chosen horizontal subsampling makes no sense iHSS=', a.ihss)
        raise SystemExit(0)
    if a.fsi3 and a.krec < 0:
        print('ERROR: chosen record to read is < 0!', a.krec)
        raise SystemExit(0)

    seeding_type = 'debug'
    if a.fmmm:
        seeding_type = 'nemoTmm'
    if a.fsi3:
        seeding_type = 'nemoTsi3'
    if a.lsidfex == 1:
        seeding_type = 'sidfex'

    ids = None
    if seeding_type in ('nemoTmm', 'nemoTsi3'):
        imaskt, xlatT, xlonT, xYt, xXt, xYf, xXf, xResKM = sit.GetModelGrid(a.fmmm)
        if a.fsi3:
            xIC = sit.GetModelSeaIceConc(a.fsi3, name=a.nsic, krec=a.krec, expected_shape=np.shape(imaskt))
        else:
            xIC = np.ones(np.shape(imaskt))
        FSmask = []
        if a.fmsk:
            FSmask = sit.GetSeedMask(a.fmsk, mvar='tmask')
            if np.shape(FSmask) != np.shape(imaskt):
                print('ERROR: `shape(FSmask) != shape(imaskt)`')
                raise SystemExit(0)
        XseedGC = sit.nemoSeed(imaskt, xlatT, xlonT, xIC, khss=a.ihss, fmsk_rstrct=FSmask)
    elif seeding_type == 'sidfex':
        XseedGC, ids = sit.SidfexSeeding(a.sidfexdat)
    else:
        XseedGC = sit.debugSeeding()
    nP = XseedGC.shape[0]
    print('\n * Shape of XseedGC =', XseedGC.shape)
    if ids is None:
        ids = np.arange(nP, dtype=int) + 1
    zTime = np.array([sit.clock2epoch(a.dat0)], dtype='i4')
    print('\n * Requested initialization date =', sit.epoch2clock(zTime[0]))
    cdate = sit.epoch2clock(zTime[0], precision='h').replace('-', '')
    XseedYX = sit.Geo2CartNPSkm1D(XseedGC)
    cextra = '_HSS' + str(a.ihss) if a.ihss > 1 else ''
    os.makedirs('./nc', exist_ok=True)
    ext = '.npz' if (a.fmmm or '').endswith('.npz') else '.nc'
    fout = './nc/sitrack_seeding_' + seeding_type + '_' + cdate + cextra + ext
    print('\n *** Saving seeding file for date =', sit.epoch2clock(zTime[0]), '\n   => into:', fout)
    sit.ncSaveCloudBuoys(fout, zTime, ids, XseedYX[None, :, 0], XseedYX[None, :, 1], XseedGC[None, :, 0],
                         XseedGC[None, :, 1], corigin='idealized_seeding')
    return fout


if __name__ == '__main__':
    main()
