"""Drop-in for `sitrack.ncio` (reference: sitrack/ncio.py).  Host-side I/O stays Python.

netCDF4 is imported lazily (it is not installed in the build image).  Every reader and
the writer also accept `.npz` files holding the SAME variable names, dtypes and leading
singleton dimensions, so the whole CLI runs on synthetic data without netCDF:
  mesh_mask : tmask glamt gphit glamf gphif glamu gphiu glamv gphiv e1t e2t
  SI3 file  : time_counter siconc u_ice v_ice
  buoy file : time id_buoy latitude longitude y_pos x_pos [mask] [time_pos]
The km coordinates of the grid come from the device projection kernel
(ConvertGeo2CartesianNPSkm -> st_latlon2xy) instead of cartopy.
"""
from os import path
from sys import argv

import numpy as np

from .util import chck4f, ConvertGeo2CartesianNPSkm
from .util import epoch2clock as e2c

__all__ = ["tunits_default", "FillValue", "GetModelGrid", "GetModelUVGrid", "GetSeedMask",
           "GetModelSeaIceConc", "ncSaveCloudBuoys", "LoadNCtime", "LoadNCdata", "SeedFileTimeInfo",
           "ModelFileTimeInfo", "open_dataset"]

tunits_default = 'seconds since 1970-01-01 00:00:00'     # ncio.py:15
FillValue = -9999.                                       # ncio.py:19


class _NpzVar:
    def __init__(self, a, units=None):
        self.a, self.units = a, units

    def __getitem__(self, k):
        return self.a[k]

    @property
    def shape(self):
        return self.a.shape


class _NpzDataset:
    """Read-only view of an .npz laid out like the netCDF files the tracker reads."""
    _time_vars = ("time", "time_counter", "time_pos")

    def __init__(self, fn):
        self.z = np.load(fn, allow_pickle=False)
        self.variables = {k: _NpzVar(self.z[k], tunits_default if k in self._time_vars else None)
                          for k in self.z.files}

        class _Dim:
            def __init__(self, n):
                self.size = n
        self.dimensions = {}
        if "time_counter" in self.z.files:
            self.dimensions["time_counter"] = _Dim(self.z["time_counter"].shape[0])
        if "time" in self.z.files:
            self.dimensions["time"] = _Dim(self.z["time"].shape[0])
        if "id_buoy" in self.z.files:
            self.dimensions["buoy"] = _Dim(self.z["id_buoy"].shape[0])

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def close(self):
        self.z.close()


def open_dataset(fn):
    """netCDF4.Dataset(fn), or the npz stand-in when `fn` ends in .npz."""
    if str(fn).endswith(".npz"):
        return _NpzDataset(fn)
    try:
        from netCDF4 import Dataset
    except ImportError as e:
        raise ImportError("netCDF4 is needed to read '%s' (or pass the .npz equivalent)" % fn) from e
    return Dataset(fn)


def _lvl(v, n):
    """v[0,..,0,:,:] with n leading singleton indices when they exist (npz files may omit them)."""
    a = np.asarray(v[:])
    while a.ndim > 2:
        a = a[0]
    return a


def GetModelGrid(fNCmeshmask, alsoF=False):
    """ncio.py:22-63 -> kmaskt, latT, lonT(0..360), Yt, Xt, Yf, Xf [km], ResKM [, kmaskf, latF, lonF]"""
    chck4f(fNCmeshmask)
    with open_dataset(fNCmeshmask) as id_mm:
        kmaskt = _lvl(id_mm.variables['tmask'], 2)
        zlonF, zlatF = _lvl(id_mm.variables['glamf'], 1), _lvl(id_mm.variables['gphif'], 1)
        zlonT, zlatT = _lvl(id_mm.variables['glamt'], 1), _lvl(id_mm.variables['gphit'], 1)
        ze1T = _lvl(id_mm.variables['e1t'], 1) / 1000.
        ze2T = _lvl(id_mm.variables['e2t'], 1) / 1000.
        if alsoF:
            kmaskf = _lvl(id_mm.variables['fmask'], 2)
    kmaskt = np.array(kmaskt, dtype='i1')
    zlonT = np.mod(zlonT, 360.)
    zlonF = np.mod(zlonF, 360.)
    zYt, zXt = ConvertGeo2CartesianNPSkm(zlatT, zlonT)
    zYf, zXf = ConvertGeo2CartesianNPSkm(zlatF, zlonF)
    zResKM = np.sqrt(ze1T * ze1T + ze2T * ze2T).astype(np.float64)
    if alsoF:
        return kmaskt, zlatT, zlonT, zYt, zXt, zYf, zXf, zResKM, kmaskf, zlatF, zlonF
    return kmaskt, zlatT, zlonT, zYt, zXt, zYf, zXf, zResKM


def GetModelUVGrid(fNCmeshmask):
    """ncio.py:66-92 -> Yv, Xv, Yu, Xu [km]"""
    chck4f(fNCmeshmask)
    with open_dataset(fNCmeshmask) as id_mm:
        zlonV, zlatV = _lvl(id_mm.variables['glamv'], 1), _lvl(id_mm.variables['gphiv'], 1)
        zlonU, zlatU = _lvl(id_mm.variables['glamu'], 1), _lvl(id_mm.variables['gphiu'], 1)
    zYv, zXv = ConvertGeo2CartesianNPSkm(zlatV, np.mod(zlonV, 360.))
    zYu, zXu = ConvertGeo2CartesianNPSkm(zlatU, np.mod(zlonU, 360.))
    return zYv, zXv, zYu, zXu


def GetSeedMask(fFSmask, mvar='tmask'):
    chck4f(fFSmask)
    with open_dataset(fFSmask) as id_mm:
        kmaskt = np.asarray(id_mm.variables[mvar][:, :])
    return np.array(kmaskt, dtype='i1')


def GetModelSeaIceConc(fNCsi3, name='siconc', krec=0, expected_shape=[]):
    chck4f(fNCsi3)
    print('    * [GetModelSeaIceConc]: reading "' + name + '" at record ' + str(krec) + ' in ' + fNCsi3 + ' !')
    with open_dataset(fNCsi3) as id_si3:
        zsic = np.asarray(id_si3.variables[name][krec, :, :])
    if len(expected_shape) > 0 and np.shape(zsic) != tuple(expected_shape):
        print('ERROR [GetModelSeaIceConc]: wrong shape for sea-ice concentration read:', np.shape(zsic),
              ', expected:', expected_shape)
        raise SystemExit(0)
    return zsic


def ncSaveCloudBuoys(cf_out, ptime, pIDs, pY, pX, pLat, pLon, mask=[], xtime=[],
                     tunits=tunits_default, fillVal=FillValue, corigin=None):
    """ncio.py:131-197: time i4, buoy i4, id_buoy i8, latitude/longitude/y_pos/x_pos f4 (time,buoy)
    [+ mask i1, time_pos i4].  Writes netCDF4 when available and the name does not end in .npz,
    otherwise an .npz with the same variables and dtypes."""
    print('\n *** [ncSaveCloudBuoys]: About to generate file: ' + cf_out + ' ...')
    (Nt,) = np.shape(ptime)
    (Nb,) = np.shape(pIDs)
    if np.shape(pY) != (Nt, Nb) or np.shape(pX) != (Nt, Nb) or np.shape(pLat) != (Nt, Nb) or np.shape(pLon) != (Nt, Nb):
        print('ERROR [ncSaveCloudBuoys]: one of the 2D arrays has a wrong shape!!!')
        raise SystemExit(0)
    lSaveMask = (np.shape(mask) == (Nt, Nb))
    lSaveTime = (np.shape(xtime) == (Nt, Nb))
    use_npz = str(cf_out).endswith(".npz")
    if not use_npz:
        try:
            from netCDF4 import Dataset
        except ImportError:
            use_npz = True
            cf_out = cf_out + ".npz"
            print('      (netCDF4 not available: writing ' + cf_out + ' with the same variables)')
    if use_npz:
        out = dict(time=np.asarray(ptime).astype('i4'), buoy=np.arange(Nb, dtype='i4'),
                   id_buoy=np.asarray(pIDs).astype('i8'), latitude=np.asarray(pLat, 'f4'),
                   longitude=np.asarray(pLon, 'f4'), y_pos=np.asarray(pY, 'f4'), x_pos=np.asarray(pX, 'f4'))
        if lSaveMask:
            out["mask"] = np.asarray(mask, 'i1')
        if lSaveTime:
            out["time_pos"] = np.asarray(xtime).astype('i4')
        np.savez_compressed(cf_out, **out)
        print('      ===> ' + cf_out + ' saved!')
        return 0
    f_out = Dataset(cf_out, 'w', format='NETCDF4')
    f_out.createDimension('time', None)
    f_out.createDimension('buoy', Nb)
    v_time = f_out.createVariable('time', 'i4', ('time',))
    v_buoy = f_out.createVariable('buoy', 'i4', ('buoy',))
    v_bid = f_out.createVariable('id_buoy', 'i8', ('buoy',))
    kw = dict(fill_value=fillVal, zlib=True, complevel=9)
    x_lat = f_out.createVariable('latitude', 'f4', ('time', 'buoy',), **kw)
    x_lon = f_out.createVariable('longitude', 'f4', ('time', 'buoy',), **kw)
    x_ykm = f_out.createVariable('y_pos', 'f4', ('time', 'buoy',), **kw)
    x_xkm = f_out.createVariable('x_pos', 'f4', ('time', 'buoy',), **kw)
    v_time.units = tunits
    v_bid.units = 'ID of buoy'
    x_lat.units = 'degrees north'
    x_lon.units = 'degrees south'          # sic (ncio.py:164)
    x_ykm.units = 'km'
    x_xkm.units = 'km'
    if lSaveMask:
        v_mask = f_out.createVariable('mask', 'i1', ('time', 'buoy',), zlib=True, complevel=9)
    if lSaveTime:
        x_tim = f_out.createVariable('time_pos', 'i4', ('time', 'buoy',), **kw)
        x_tim.units = tunits
    v_buoy[:] = np.arange(Nb, dtype='i8')
    v_bid[:] = pIDs[:]
    for jt in range(Nt):
        v_time[jt] = ptime[jt]
        x_lat[jt, :] = pLat[jt, :]
        x_lon[jt, :] = pLon[jt, :]
        x_ykm[jt, :] = pY[jt, :]
        x_xkm[jt, :] = pX[jt, :]
        if lSaveMask:
            v_mask[jt, :] = mask[jt, :]
        if lSaveTime:
            x_tim[jt, :] = xtime[jt, :]
    if corigin:
        f_out.Origin = corigin
    f_out.About = 'Lagrangian sea-ice drift'
    f_out.Author = 'Generated with `' + path.basename(argv[0]) + '` of `sitrack` (L. Brodeau, 2023)'
    f_out.close()
    print('      ===> ' + cf_out + ' saved!')
    return 0


def _check_tunits(v, who):
    u = getattr(v, "units", tunits_default)
    if u is not None and u != tunits_default:
        print(' ERROR [' + who + '()]: we expect "' + tunits_default + '" as units for the time record vector, yet we have: ' + str(u))
        raise SystemExit(0)


def LoadNCtime(cfile, ltime2d=False, iverbose=0):
    """ncio.py:199-239 -> Nt, time [, time_pos]"""
    chck4f(cfile)
    with open_dataset(cfile) as id_in:
        if 'time' not in id_in.dimensions or 'time' not in id_in.variables:
            print(' ERROR [LoadNCtime()]: no `time` found into input file!')
            raise SystemExit(0)
        Nt = id_in.dimensions['time'].size
        _check_tunits(id_in.variables['time'], 'LoadNCtime')
        print('    * [LoadNCtime] => reading "time" (' + str(Nt) + ' records) in file ' + path.basename(cfile))
        ztime = np.asarray(id_in.variables['time'][:])
        if ltime2d:
            if 'time_pos' not in id_in.variables:
                print(' ERROR [LoadNCtime()]: no variable `time_pos` found into input file!')
                raise SystemExit(0)
            _check_tunits(id_in.variables['time_pos'], 'LoadNCtime')
            ztime2d = np.asarray(id_in.variables['time_pos'][:, :])
            if ztime2d.shape[0] != Nt:
                print(' ERROR [LoadNCtime()]: array `time_pos` has not the same number of records as `time`!!!')
                raise SystemExit(0)
            return Nt, ztime, ztime2d
        return Nt, ztime


def LoadNCdata(cfile, krec=-1, lmask=False, lGetTimePos=False, iverbose=0):
    """ncio.py:243-326 -> time, IDs, LatLon (..,nP,2) with lon in 0..360, YX (..,nP,2) [, mask][, time_pos]"""
    need = ['id_buoy', 'latitude', 'longitude', 'y_pos', 'x_pos'] + (['time_pos'] if lGetTimePos else [])
    chck4f(cfile)
    with open_dataset(cfile) as id_in:
        for cd in ['time', 'buoy']:
            if cd not in id_in.dimensions:
                print(' ERROR [LoadNCdata()]: no dimensions `' + cd + '` found into input file!')
                raise SystemExit(0)
        for cv in need:
            if cv not in id_in.variables:
                print(' ERROR [LoadNCdata()]: no variable `' + cv + '` found into input file!')
                raise SystemExit(0)
        Nt = id_in.dimensions['time'].size
        nP = id_in.dimensions['buoy'].size
        _check_tunits(id_in.variables['time'], 'LoadNCdata')
        idxR = krec if krec >= 0 else np.arange(Nt, dtype=int)
        ztime = np.asarray(id_in.variables['time'][idxR])
        kBIDs = np.zeros(nP, dtype=int)
        kBIDs[:] = id_in.variables['id_buoy'][:]
        zlat = np.asarray(id_in.variables['latitude'][idxR, :])
        zlon = np.array(id_in.variables['longitude'][idxR, :])
        zy = np.asarray(id_in.variables['y_pos'][idxR, :])
        zx = np.asarray(id_in.variables['x_pos'][idxR, :])
        if lmask:
            zmsk = np.asarray(id_in.variables['mask'][idxR, :])
        if lGetTimePos:
            ztpos = np.asarray(id_in.variables['time_pos'][idxR, :])
    zlon[:] = np.mod(zlon, 360.)
    shp = (nP, 2) if krec >= 0 else (Nt, nP, 2)
    zLatLon, zYX = np.zeros(shp), np.zeros(shp)
    zLatLon[..., 0], zLatLon[..., 1] = zlat, zlon            # f4 -> f8 like the reference
    zYX[..., 0], zYX[..., 1] = zy, zx
    out = [ztime, kBIDs, zLatLon, zYX]
    if lmask:
        out.append(zmsk)
    if lGetTimePos:
        out.append(ztpos)
    return tuple(out)


def SeedFileTimeInfo(fSeedNc, ltime2d=False, iverbose=0):
    """ncio.py:329-352 -> idate0, idateN (rounded to the hour), SeedName, SeedBatch, time_pos"""
    from re import split
    from math import ceil, floor
    base = path.basename(fSeedNc)
    cSeed = base.replace('SELECTION_', '').replace('.npz', '').replace('.nc', '')
    cBtch = split('_', base)[2]
    chck4f(fSeedNc)
    if ltime2d:
        ntr, zt, zt2d = LoadNCtime(fSeedNc, ltime2d=True, iverbose=iverbose)
        idate0, idateN = np.min(zt2d), np.max(zt2d)
    else:
        ntr, zt = LoadNCtime(fSeedNc, iverbose=iverbose)
        idate0, idateN = zt[0], zt[ntr - 1]
        zt2d = []
    print('    * [SeedFileTimeInfo] => earliest and latest time position in the SEED file: ' + e2c(idate0) + ' - ' + e2c(idateN))
    idate0, idateN = int(floor(idate0 / 3600.) * 3600.), int(ceil(idateN / 3600.) * 3600.)
    print('    * [SeedFileTimeInfo]  ==> will actually use rounded to the hour! => ' + e2c(idate0) + ' - ' + e2c(idateN))
    return idate0, idateN, cSeed, cBtch, zt2d


def ModelFileTimeInfo(fModelNc, iverbose=0):
    """ncio.py:356-384 -> Nt, time_counter (i4), idate0, idateN, CONF, EXP (from the FILE NAME)"""
    from re import split
    with open_dataset(fModelNc) as ds_mod:
        Nt = ds_mod.dimensions['time_counter'].size
        _check_tunits(ds_mod.variables['time_counter'], 'ModelFileTimeInfo')
        ztime = np.array(ds_mod.variables['time_counter'][:], dtype='i4')
    print('    * [ModelFileTimeInfo] => ' + str(Nt) + ' records in input MODEL file!')
    idate0, idateN = np.min(ztime), np.max(ztime)
    print('    * [ModelFileTimeInfo] => earliest and latest time position in the MODEL file: ' + e2c(idate0) + ' - ' + e2c(idateN))
    vn = split('_', path.basename(fModelNc))
    zz = split('-', vn[1])
    nconf = vn[0]
    if len(zz) == 1:
        zz = split('-', vn[0])
        nconf = zz[0]
    nexpr = zz[1]
    print('    * [ModelFileTimeInfo] => NEMO config and experiment =', nconf, nexpr, '\n')
    return Nt, ztime, idate0, idateN, nconf, nexpr
