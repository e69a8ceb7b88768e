/*
 * st_oracle.c -- CPU restatement of sitrack's buoy-advection hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under sitrack_b200/ may include, link
 * or call this file; it is the checker for tests/, __graft_entry__.smoke()
 * and the cpu_baseline leg of bench.py, never the thing shipped or measured
 * as the product.
 *
 * Every function states the reference lines (relative to the upstream repo
 * stephanieleroux/sitrack) whose arithmetic it follows.  The parity-critical
 * arithmetic is IEEE double + - * / and comparisons in the order the Python
 * expressions evaluate them; compile with -ffp-contract=off so gcc never
 * fuses a*b+c (Python never does).
 *
 * Pinning (see tests/golden/ and oracle/make_golden.py):
 *   - geometry predicates, the one-hop cell walk, Survive, SeedInit and the
 *     record x buoy loop are pinned against outputs of the reference's own
 *     Python functions imported in the build container;
 *   - the polar-stereographic inverse (orc_inv_stere) restates PROJ's
 *     `stere` (reached by the reference through cartopy -> pyproj, neither
 *     present nor version-pinned upstream): PARITY UNPINNED for lat/lon.
 *
 * Array conventions: all 2-D grids are row-major (Nj, Ni); points are
 * [y, x] pairs in km; indices are (j, i).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_FILL (-9999.0)            /* sitrack/ncio.py:19 FillValue */

typedef struct { double y, x; } pt_t;

static inline pt_t P2(const double *Y, const double *X, int Ni, long j, long i)
{
    pt_t p; p.y = Y[j * (long)Ni + i]; p.x = X[j * (long)Ni + i]; return p;
}

/* ---- sitrack/tracking.py:44-49  _ccw_ ---------------------------------- */
static inline int orc_ccw(pt_t A, pt_t B, pt_t C)
{
    return (C.y - A.y) * (B.x - A.x) > (B.y - A.y) * (C.x - A.x);
}

/* ---- sitrack/tracking.py:51-58  intersect2Seg -------------------------- */
int orc_intersect2seg(const double *A, const double *B, const double *C, const double *D)
{
    pt_t a = {A[0], A[1]}, b = {B[0], B[1]}, c = {C[0], C[1]}, d = {D[0], D[1]};
    return (orc_ccw(a, c, d) != orc_ccw(b, c, d)) && (orc_ccw(a, b, c) != orc_ccw(a, b, d));
}
static inline int isect(pt_t a, pt_t b, pt_t c, pt_t d)
{
    return (orc_ccw(a, c, d) != orc_ccw(b, c, d)) && (orc_ccw(a, b, c) != orc_ccw(a, b, d));
}

/* ---- sitrack/locate.py:49-78  IsInsideQuadrangle -----------------------
 * quad is 4 x [y,x].  The loop runs n+1 = 5 times with vertex quad[i%4];
 * pass 0 compares vertex 0 with itself and can never toggle.             */
int orc_inside_quad(double y, double x, const double *quad)
{
    int inside = 0;
    double xints = 0.0;
    double y1 = quad[0], x1 = quad[1];
    for (int i = 0; i < 5; ++i) {
        double y2 = quad[2 * (i % 4)], x2 = quad[2 * (i % 4) + 1];
        double ymin = (y2 < y1) ? y2 : y1;      /* Python min(a,b): b if b<a else a */
        double ymax = (y2 > y1) ? y2 : y1;
        double xmax = (x2 > x1) ? x2 : x1;
        if (y > ymin) {
            if (y <= ymax) {
                if (x <= xmax) {
                    if (y1 != y2)
                        xints = (y - y1) * (x2 - x1) / (y2 - y1) + x1;
                    if (x1 == x2 || x <= xints)
                        inside = !inside;
                }
            }
        }
        y1 = y2; x1 = x2;
    }
    return inside;
}

/* ---- sitrack/tracking.py:62-93  Survive --------------------------------
 * ic == NULL reproduces nothing sensible upstream (UnboundLocalError), so
 * callers must pass a concentration field; ic is the f8 copy of the f4
 * record (si3_part_tracker.py:372) -- here converted on the fly.          */
int orc_survive(long jT, long iT, int Nj, int Ni, const int8_t *tmask, const float *ic,
                double rmin_conc)
{
    if (jT == 0 || jT == 1 || jT == Nj - 2 || jT == Nj - 1 ||
        iT == 0 || iT == 1 || iT == Ni - 2 || iT == Ni - 1)
        return 1;
#define AT(a, j, i) ((a)[(j) * (long)Ni + (i)])
    int zmt = AT(tmask, jT, iT) + AT(tmask, jT, iT + 1) + AT(tmask, jT + 1, iT)
            + AT(tmask, jT, iT - 1) + AT(tmask, jT - 1, iT - 1);   /* sic: [jT-1,iT-1] */
    if (zmt < 5) return 1;
    double zic = 0.2 * ((double)AT(ic, jT, iT) + (double)AT(ic, jT, iT + 1) + (double)AT(ic, jT + 1, iT)
                        + (double)AT(ic, jT, iT - 1) + (double)AT(ic, jT - 1, iT - 1));
#undef AT
    if (zic < rmin_conc) return 1;
    return 0;
}

/* vertices of the host cell of T[jT,iT], anti-clockwise from bottom-left:
 * sitrack/locate.py:320-321                                               */
static inline void cell_vertices(long jT, long iT, long jv[4], long iv[4])
{
    jv[0] = jT - 1; jv[1] = jT - 1; jv[2] = jT; jv[3] = jT;
    iv[0] = iT - 1; iv[1] = iT;     iv[2] = iT; iv[3] = iT - 1;
}

/* ---- sitrack/tracking.py:182-200  CrossedEdge --------------------------
 * First of bottom/right/top/left whose edge meets P1->P2; when none does
 * the Python loop variable is left at 3, so the answer is 4.              */
int orc_crossed_edge(pt_t p1, pt_t p2, const long jv[4], const long iv[4],
                     const double *Y, const double *X, int Ni)
{
    int kk;
    for (kk = 0; kk < 4; ++kk) {
        int kp1 = (kk + 1) % 4;
        if (isect(p1, p2, P2(Y, X, Ni, jv[kk], iv[kk]), P2(Y, X, Ni, jv[kp1], iv[kp1])))
            return kk + 1;
    }
    return 4;
}

/* ---- sitrack/tracking.py:203-249  NewHostCell -------------------------- */
int orc_new_host_cell(int kcross, pt_t p1, pt_t p2, const long jv[4], const long iv[4],
                      const double *Y, const double *X, int Ni)
{
    long jbl = jv[0], jbr = jv[1], jur = jv[2], jul = jv[3];
    long ibl = iv[0], ibr = iv[1], iur = iv[2], iul = iv[3];
    int k = kcross;
    if (kcross == 1) {
        if      (isect(p1, p2, P2(Y, X, Ni, jbl, ibl), P2(Y, X, Ni, jbl - 1, ibl))) k = 5;
        else if (isect(p1, p2, P2(Y, X, Ni, jbr, ibr), P2(Y, X, Ni, jbr - 1, ibr))) k = 6;
    } else if (kcross == 2) {
        if      (isect(p1, p2, P2(Y, X, Ni, jbr, ibr), P2(Y, X, Ni, jbr, ibr + 1))) k = 6;
        else if (isect(p1, p2, P2(Y, X, Ni, jur, iur), P2(Y, X, Ni, jur, iur + 1))) k = 7;
    } else if (kcross == 3) {
        if      (isect(p1, p2, P2(Y, X, Ni, jul, iul), P2(Y, X, Ni, jul + 1, iul))) k = 8;
        else if (isect(p1, p2, P2(Y, X, Ni, jur, iur), P2(Y, X, Ni, jur + 1, iur))) k = 7;
    } else if (kcross == 4) {
        if      (isect(p1, p2, P2(Y, X, Ni, jul, iul), P2(Y, X, Ni, jul, iul - 1))) k = 8;
        else if (isect(p1, p2, P2(Y, X, Ni, jbl, ibl), P2(Y, X, Ni, jbl, ibl - 1))) k = 5;
    }
    return k;
}

/* ---- sitrack/tracking.py:253-305  UpdtInd4NewCell (index shift only) --- */
int orc_updt_ind(int knhc, long *jT, long *iT)
{
    switch (knhc) {
    case 1: *jT -= 1;            break;
    case 2:            *iT += 1; break;
    case 3: *jT += 1;            break;
    case 4:            *iT -= 1; break;
    case 5: *jT -= 1;  *iT -= 1; break;
    case 6: *jT -= 1;  *iT += 1; break;
    case 7: *jT += 1;  *iT += 1; break;
    case 8: *jT += 1;  *iT -= 1; break;
    default: return -1;
    }
    return 0;
}

/* thin exported wrappers so the Python tests can pin each predicate ------ */
int orc_crossed_edge_c(const double *p1, const double *p2, long jT, long iT,
                       const double *Y, const double *X, int Ni)
{
    long jv[4], iv[4]; cell_vertices(jT, iT, jv, iv);
    pt_t a = {p1[0], p1[1]}, b = {p2[0], p2[1]};
    return orc_crossed_edge(a, b, jv, iv, Y, X, Ni);
}
int orc_new_host_cell_c(int kcross, const double *p1, const double *p2, long jT, long iT,
                        const double *Y, const double *X, int Ni)
{
    long jv[4], iv[4]; cell_vertices(jT, iT, jv, iv);
    pt_t a = {p1[0], p1[1]}, b = {p2[0], p2[1]};
    return orc_new_host_cell(kcross, a, b, jv, iv, Y, X, Ni);
}

/* ======================================================================= *
 *  Polar stereographic, ellipsoidal (PROJ "stere", north-pole branch).
 *  Reference reaches it via sitrack/util.py:413-429 -> cartopy
 *  NorthPolarStereo(central_longitude=-45, true_scale_latitude=70) on the
 *  default WGS84 globe.  PARITY UNPINNED (PROJ is not under the reference).
 * ======================================================================= */
#define WGS84_A   6378137.0
#define WGS84_F   (1.0 / 298.257223563)
#define HALFPI    1.5707963267948966
#define PI_       3.14159265358979323846
#define DEG2RAD   0.017453292519943295
#define RAD2DEG   57.29577951308232

static double tsfn(double phi, double sinphi, double e)
{
    double es = e * sinphi;
    return tan(0.5 * (HALFPI - phi)) / pow((1.0 - es) / (1.0 + es), 0.5 * e);
}
static double stere_akm1(double lat_ts_deg, double e)
{
    double phits = fabs(lat_ts_deg) * DEG2RAD;
    if (fabs(phits - HALFPI) < 1e-10)
        return 2.0 / sqrt(pow(1.0 + e, 1.0 + e) * pow(1.0 - e, 1.0 - e));
    double t = sin(phits);
    double akm1 = cos(phits) / tsfn(phits, t, e);
    t *= e;
    return akm1 / sqrt(1.0 - t * t);
}
static double adjlon(double lam)
{
    if (fabs(lam) <= PI_) return lam;
    lam += PI_;
    lam -= 2.0 * PI_ * floor(lam / (2.0 * PI_));
    return lam - PI_;
}

/* km (y,x) -> degrees (lat,lon), n points, yx and latlon are (n,2), on the ellipsoid (a [m], f). */
void orc_inv_stere_ell(long n, const double *yx, double *latlon, double lat_ts, double lon0, double a_m, double f)
{
    const double es = f * (2.0 - f), e = sqrt(es);
    const double akm1 = stere_akm1(lat_ts, e);
    for (long k = 0; k < n; ++k) {
        double x = (1000.0 * yx[2 * k + 1]) / a_m;
        double y = (1000.0 * yx[2 * k + 0]) / a_m;
        double rho = hypot(x, y);
        y = -y;                                       /* N_POLE */
        double tp = -rho / akm1;
        double phi_l = HALFPI - 2.0 * atan(tp);
        const double halfpi = -HALFPI, halfe = -0.5 * e;
        double phi = phi_l, lam = 0.0;
        for (int i = 8; i--; phi_l = phi) {
            double sinphi = e * sin(phi_l);
            phi = 2.0 * atan(tp * pow((1.0 + sinphi) / (1.0 - sinphi), halfe)) - halfpi;
            if (fabs(phi_l - phi) < 1e-10) break;
        }
        lam = (x == 0.0 && y == 0.0) ? 0.0 : atan2(x, y);
        lam = adjlon(lam + lon0 * DEG2RAD);
        latlon[2 * k + 0] = phi * RAD2DEG;
        latlon[2 * k + 1] = lam * RAD2DEG;
    }
}

/* the reference's case: WGS84, the globe cartopy gives NorthPolarStereo by default */
void orc_inv_stere(long n, const double *yx, double *latlon, double lat_ts, double lon0)
{
    orc_inv_stere_ell(n, yx, latlon, lat_ts, lon0, WGS84_A, WGS84_F);
}

/* degrees (lat,lon) -> km (y,x) ; sitrack/util.py:394-410,434-451 */
void orc_fwd_stere_ell(long n, const double *latlon, double *yx, double lat_ts, double lon0, double a_m, double f)
{
    const double es = f * (2.0 - f), e = sqrt(es);
    const double akm1 = stere_akm1(lat_ts, e);
    for (long k = 0; k < n; ++k) {
        double phi = latlon[2 * k + 0] * DEG2RAD;
        double lam = adjlon(latlon[2 * k + 1] * DEG2RAD - lon0 * DEG2RAD);
        double sinphi = sin(phi);
        double x = (fabs(phi - HALFPI) < 1e-15) ? 0.0 : akm1 * tsfn(phi, sinphi, e);
        double y = -x * cos(lam);                     /* N_POLE: y = -rho*cos(lam) */
        x = x * sin(lam);
        yx[2 * k + 0] = a_m * y / 1000.0;
        yx[2 * k + 1] = a_m * x / 1000.0;
    }
}
void orc_fwd_stere(long n, const double *latlon, double *yx, double lat_ts, double lon0)
{
    orc_fwd_stere_ell(n, latlon, yx, lat_ts, lon0, WGS84_A, WGS84_F);
}

/* ======================================================================= *
 *  Seeding: NearestPoint + Survive + FindContainingCell
 * ======================================================================= */

/* sitrack/util.py:85-103  Haversine (scalar form of the vector expression) */
double orc_haversine(double plat, double plon, double xlat, double xlon)
{
    const double to_rad = 3.141592653589793 / 180.;
    const double R = 6360.;
    double a1 = sin(0.5 * ((xlat - plat) * to_rad));
    double a2 = sin(0.5 * ((xlon - plon) * to_rad));
    double a3 = cos(xlat * to_rad) * cos(plat * to_rad);
    return 2. * R * asin(sqrt(a1 * a1 + a3 * a2 * a2));
}

/* sitrack/locate.py:222-276 NearestPoint as SeedInit calls it
 * (tracking.py:134: rd_found_km=2.5, resolkm 2-D, no ji_prv box, max_itr=10).
 * Whole-grid scan, first-minimum argmin in C order (locate.py:13-20).
 * With a 2-D resolkm the rd_found_km argument is overwritten on pass 1.   */
void orc_nearest_point(double latP, double lonP, int Nj, int Ni,
                       const double *lat, const double *lon, const double *reskm,
                       double rd_found_km, int max_itr, long *jy, long *jx, double *dmin_out)
{
    long n = (long)Nj * Ni, kmin = 0;
    double dmin = INFINITY;
    for (long k = 0; k < n; ++k) {
        double d = orc_haversine(latP, lonP, lat[k], lon[k]);
        if (d < dmin) { dmin = d; kmin = k; }
    }
    long j = kmin / Ni, i = kmin % Ni;
    double rfnd = rd_found_km;
    int igo = 0, lfound = 0;
    while (!lfound && igo < max_itr) {
        igo += 1;
        if (igo == 1 && reskm) rfnd = 0.5 * reskm[j * (long)Ni + i];
        if (igo == 1) igo = 2;                         /* no box: skip one round */
        lfound = (dmin < rfnd);
        if (igo > 1 && !lfound) rfnd = 1.2 * rfnd;
    }
    if (igo == max_itr) { j = -1; i = -1; }
    *jy = j; *jx = i;
    if (dmin_out) *dmin_out = dmin;
}

/* sitrack/locate.py:280-330 FindContainingCell; returns lPin, cell in jT,iT */
int orc_find_containing_cell(double y, double x, long kj, long ki,
                             const double *Yf, const double *Xf, int Ni, long *jT, long *iT)
{
    static const int dj[5] = {0, 0, 1, 0, -1};
    static const int di[5] = {0, 1, 0, -1, 0};
    int lPin = 0;
    long j = kj, i = ki;
    for (int kp = 0; kp < 5 && !lPin; ++kp) {
        j = kj + dj[kp]; i = ki + di[kp];
        long jv[4], iv[4]; cell_vertices(j, i, jv, iv);
        double quad[8];
        for (int v = 0; v < 4; ++v) {
            pt_t p = P2(Yf, Xf, Ni, jv[v], iv[v]);
            quad[2 * v] = p.y; quad[2 * v + 1] = p.x;
        }
        lPin = orc_inside_quad(y, x, quad);
    }
    *jT = j; *iT = i;
    return lPin;
}

/* sitrack/tracking.py:98-178 SeedInit: per buoy NearestPoint -> Survive ->
 * FindContainingCell; keep[] is the kmask, jiT (nP,2) the containing cell.
 * ic0 = siconc record at kstrt (si3_part_tracker.py:228-229).             */
void orc_seed_init(long nP, const double *SG /*(nP,2) lat,lon*/, const double *SC /*(nP,2) y,x*/,
                   int Nj, int Ni, const double *latT, const double *lonT,
                   const double *Yf, const double *Xf, const double *reskm,
                   const int8_t *tmask, const float *ic0, double rmin_conc,
                   int64_t *jiT, int8_t *keep, int64_t *jiNearest)
{
#pragma omp parallel for schedule(dynamic, 4)
    for (long p = 0; p < nP; ++p) {
        long jT, iT;
        keep[p] = 1; jiT[2 * p] = 0; jiT[2 * p + 1] = 0;
        orc_nearest_point(SG[2 * p], SG[2 * p + 1], Nj, Ni, latT, lonT, reskm, 2.5, 10, &jT, &iT, 0);
        if (jiNearest) { jiNearest[2 * p] = jT; jiNearest[2 * p + 1] = iT; }
        if (jT < 0 || iT < 0) keep[p] = 0;
        if (keep[p] && orc_survive(jT, iT, Nj, Ni, tmask, ic0, rmin_conc) > 0) keep[p] = 0;
        if (keep[p]) {
            long jc, ic_;
            int lPin = orc_find_containing_cell(SC[2 * p], SC[2 * p + 1], jT, iT, Yf, Xf, Ni, &jc, &ic_);
            jiT[2 * p] = jc; jiT[2 * p + 1] = ic_;
            if (!lPin) keep[p] = 0;
        }
    }
}

/* ======================================================================= *
 *  The record x buoy loop, si3_part_tracker.py:361-496
 *
 *  posC (nrec+1, nP, 2) [y,x] km is both state and output: the caller
 *  fills row 0 (or row k0 per buoy without -F, :335-344) and ORC_FILL
 *  elsewhere.  posG (nrec+1, nP, 2) [lat,lon] rows >=1 are recomputed for
 *  every buoy each record (:493), fill rows included.  mask (nrec+1, nP).
 *  jiT (nP,2) and alive (nP) are updated in place.  Optional histories:
 *  jiT_hist (nrec+1,nP,2) / alive_hist (nrec+1,nP) hold the state AFTER
 *  each record (row 0 = initial).  U,V,IC: (nrec, Nj, Ni) f4, record jt of
 *  the stack is file record jrec = jt + kstrt.
 * ======================================================================= */
long orc_track(int Nj, int Ni,
               const double *Yf, const double *Xf, const double *Yu, const double *Xu,
               const double *Yv, const double *Xv, const int8_t *tmask,
               int nrec, int kstrt, const float *U, const float *V, const float *IC,
               long nP, double *posC, double *posG, int8_t *mask,
               int64_t *jiT, int8_t *alive,
               const int32_t *rec_first, const int32_t *rec_last,
               int uv_strategy, double rdt, double rmin_conc, int do_latlon,
               int64_t *nalive_rec, int32_t *jiT_hist, int8_t *alive_hist)
{
    long ncross = 0;
    const long npt = (long)Nj * Ni;
    if (jiT_hist) for (long p = 0; p < nP; ++p) {
        jiT_hist[2 * p] = (int32_t)jiT[2 * p]; jiT_hist[2 * p + 1] = (int32_t)jiT[2 * p + 1];
    }
    if (alive_hist) memcpy(alive_hist, alive, (size_t)nP);

    for (int jt = 0; jt < nrec; ++jt) {
        const int jrec = jt + kstrt;
        const float *xU = U + (long)jt * npt, *xV = V + (long)jt * npt, *xIC = IC + (long)jt * npt;
        double *cur = posC + (long)jt * nP * 2, *nxt = posC + (long)(jt + 1) * nP * 2;
        int8_t *mnxt = mask + (long)(jt + 1) * nP;
        long na = 0;
        for (long p = 0; p < nP; ++p) na += alive[p];
        if (nalive_rec) nalive_rec[jt] = na;

#pragma omp parallel for schedule(static) reduction(+:ncross)
        for (long p = 0; p < nP; ++p) {
            int f = rec_first ? rec_first[p] : kstrt;
            int l = rec_last ? rec_last[p] : kstrt + nrec - 1;
            if (!(alive[p] == 1 && jrec >= f && jrec <= l)) continue;

            pt_t P = {cur[2 * p], cur[2 * p + 1]};
            long jT = jiT[2 * p], iT = jiT[2 * p + 1];
            double zU, zV;
#define FLD(a, j, i) ((double)(a)[(j) * (long)Ni + (i)])
            if (uv_strategy == 0) {                       /* :423-425 */
                zU = 0.5 * (FLD(xU, jT, iT) + FLD(xU, jT, iT - 1));
                zV = 0.5 * (FLD(xV, jT, iT) + FLD(xV, jT - 1, iT));
            } else {                                      /* :427-441 */
                pt_t Fp = P2(Yf, Xf, Ni, jT, iT);
                int llum1 = isect(P, Fp, P2(Yv, Xv, Ni, jT - 1, iT), P2(Yv, Xv, Ni, jT, iT));
                int llvm1 = isect(P, Fp, P2(Yu, Xu, Ni, jT, iT - 1), P2(Yu, Xu, Ni, jT, iT));
                zU = llum1 ? FLD(xU, jT, iT - 1) : FLD(xU, jT, iT);
                zV = llvm1 ? FLD(xV, jT - 1, iT) : FLD(xV, jT, iT);
            }
#undef FLD
            double dx = zU * rdt, dy = zV * rdt;          /* :452-453 */
            pt_t Pn;
            Pn.x = P.x + dx / 1000.;                      /* :457 */
            Pn.y = P.y + dy / 1000.;                      /* :458 */
            nxt[2 * p] = Pn.y; nxt[2 * p + 1] = Pn.x;     /* :459 */
            mnxt[p] = 1;                                  /* :460 */

            long jv[4], iv[4]; cell_vertices(jT, iT, jv, iv);
            double quad[8];
            for (int v = 0; v < 4; ++v) {
                pt_t q = P2(Yf, Xf, Ni, jv[v], iv[v]);
                quad[2 * v] = q.y; quad[2 * v + 1] = q.x;
            }
            if (!orc_inside_quad(Pn.y, Pn.x, quad)) {     /* :466-484 */
                int icross = orc_crossed_edge(P, Pn, jv, iv, Yf, Xf, Ni);
                int inhc = orc_new_host_cell(icross, P, Pn, jv, iv, Yf, Xf, Ni);
                orc_updt_ind(inhc, &jT, &iT);
                jiT[2 * p] = jT; jiT[2 * p + 1] = iT;
                if (orc_survive(jT, iT, Nj, Ni, tmask, xIC, rmin_conc) > 0) alive[p] = 0;
                ncross += 1;
            }
        }
        if (do_latlon)                                    /* :493, every row */
            orc_inv_stere(nP, nxt, posG + (long)(jt + 1) * nP * 2, 70., -45.);
        if (jiT_hist) {
            int32_t *h = jiT_hist + (long)(jt + 1) * nP * 2;
            for (long p = 0; p < nP; ++p) { h[2 * p] = (int32_t)jiT[2 * p]; h[2 * p + 1] = (int32_t)jiT[2 * p + 1]; }
        }
        if (alive_hist) memcpy(alive_hist + (long)(jt + 1) * nP, alive, (size_t)nP);
    }
    return ncross;
}
