#!/usr/bin/env python3
"""Write a synthetic NANUK4-shaped case as the three input files the tracker's CLI expects
(.npz stand-ins of the netCDF files, same variable names / dtypes / file-name conventions):

    python tools/make_synth_case.py OUTDIR [--grid nanuk4|small|tiny] [--nrec 24] [--hss 5]

  OUTDIR/mesh_mask_<CONF>.npz                       tmask glam[tfuv] gphi[tfuv] e1t e2t
  OUTDIR/<CONF>-SYN00_1h_<d0>_<d1>_icemod.npz       time_counter siconc u_ice v_ice
  OUTDIR/sitrack_seeding_nemoTsi3_<d0>_00_HSS<n>.npz  time id_buoy latitude longitude y_pos x_pos

then:  python si3_part_tracker.py -i <icemod> -m <mesh_mask> -s <seeding> -F -N <CONF>
"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import synth                                    # noqa: E402
from synth.grid import km_to_latlon             # noqa: E402
from synth.records import T0_EPOCH, time_counter    # noqa: E402


def write_case(outdir, grid="nanuk4", nrec=24, hss=5, conf="SYNTH4", seed=0):
    os.makedirs(outdir, exist_ok=True)
    g = synth.make_grid(**synth.GRID_PRESETS[grid], seed=seed)
    U, V, IC = synth.make_records(g, nrec, seed=seed + 1)
    mm = dict(tmask=g["tmask"][None, None].astype("i1"))
    for pt, (Y, X) in dict(t=("Yt", "Xt"), f=("Yf", "Xf"), u=("Yu", "Xu"), v=("Yv", "Xv")).items():
        lat, lon = km_to_latlon(g[Y], g[X])
        mm["gphi" + pt], mm["glam" + pt] = lat[None], lon[None]
    e1 = np.hypot(np.diff(g["Yu"], axis=1, prepend=np.nan), np.diff(g["Xu"], axis=1, prepend=np.nan)); e1[:, 0] = e1[:, 1]
    e2 = np.hypot(np.diff(g["Yv"], axis=0, prepend=np.nan), np.diff(g["Xv"], axis=0, prepend=np.nan)); e2[0, :] = e2[1, :]
    mm["e1t"], mm["e2t"] = (1000.0 * e1)[None], (1000.0 * e2)[None]
    f_mm = os.path.join(outdir, "mesh_mask_%s.npz" % conf)
    np.savez(f_mm, **mm)
    f_si3 = os.path.join(outdir, "%s-SYN00_1h_19961215_19961216_icemod.npz" % conf)
    np.savez(f_si3, time_counter=time_counter(nrec), siconc=IC, u_ice=U, v_ice=V)
    ids, SG, SC = synth.hss_seeds(g, IC[0], khss=hss)
    lon = np.where(SG[:, 1] > 180.0, SG[:, 1] - 360.0, SG[:, 1])
    f_seed = os.path.join(outdir, "sitrack_seeding_nemoTsi3_19961215_00_HSS%d.npz" % hss)
    np.savez(f_seed, time=np.array([T0_EPOCH], "i4"), id_buoy=ids.astype("i8"),
             latitude=SG[:, 0].astype("f4")[None], longitude=lon.astype("f4")[None],
             y_pos=SC[:, 0].astype("f4")[None], x_pos=SC[:, 1].astype("f4")[None])
    return dict(mesh=f_mm, si3=f_si3, seed=f_seed, grid=g, records=(U, V, IC), conf=conf)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("outdir")
    ap.add_argument("--grid", default="nanuk4", choices=list(synth.GRID_PRESETS))
    ap.add_argument("--nrec", type=int, default=24)
    ap.add_argument("--hss", type=int, default=5)
    a = ap.parse_args()
    r = write_case(a.outdir, a.grid, a.nrec, a.hss)
    print("python si3_part_tracker.py -i %s -m %s -s %s -F -N %s" % (r["si3"], r["mesh"], r["seed"], r["conf"]))
