// st_device.cuh -- device-side geometry and projection primitives of the
// buoy-advection hot path (sm_100a).  One source of truth for every kernel.
//
// Parity rule: every quantity the reference computes with Python floats is
// rebuilt here from the SAME IEEE-754 double operations in the SAME order,
// spelled with __dadd_rn/__dsub_rn/__dmul_rn/__ddiv_rn so that ptxas can never
// contract a*b+c into an FMA (Python cannot), whatever -fmad says.  Reference
// lines are cited per function (paths relative to stephanieleroux/sitrack).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace st {

#define ST_FILL (-9999.0)                 // sitrack/ncio.py:19 FillValue
// A discontinued buoy (iAlive = 0, si3_part_tracker.py:483-484) carries bit 31 in the jT word of its cell as well
// as alive = 0: the default step kernel streams 24 B of state per buoy (pos, cell) and never touches `alive`.
#define ST_DEAD_BIT ((int)0x80000000)

// A point of the km plane, stored [y, x] like every coordinate pair upstream.
struct __align__(16) pt { double y, x; };

// Geometry gathers carry an L2 evict_last policy: ~1 GB of touch-once state and rows streams through
// the 126 MB L2 every launch, and without the hint the cell geometry (re-used by the next record) is
// evicted and every warp's first gather pays an HBM round trip.
__device__ __forceinline__ unsigned long long l2_keep_policy()
{
    unsigned long long pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// Debug builds (python -m sitrack_b200.build -DST_DEBUG_BOUNDS=1 -o lib_dbg.so; SITRACK_B200_LIB=lib_dbg.so):
// the step kernels check every cell whose stencil they are about to gather against the grid and trap on a
// violation (compute-sanitizer is not available on every pool).
#ifdef ST_DEBUG_BOUNDS
#define ST_CHECK_CELL(c, reach, Nj, Ni) do { const long long c_ = (c), r_ = (reach), n_ = (long long)(Nj) * (Ni); \
                                             if (c_ - r_ < 0 || c_ + r_ >= n_) __trap(); } while (0)
#else
#define ST_CHECK_CELL(c, reach, Nj, Ni) do { } while (0)
#endif
__device__ __forceinline__ pt ldg_pt(const pt* __restrict__ a, int idx)
{
    pt r;
    asm("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;"
        : "=d"(r.y), "=d"(r.x) : "l"(a + idx), "l"(l2_keep_policy()));
    return r;
}
// streaming (touch-once) accesses: keep them out of the way of the resident
// geometry / velocity record in L1 and L2
__device__ __forceinline__ pt ld_stream_pt(const pt* a)
{
    const double2 v = __ldcs(reinterpret_cast<const double2*>(a));
    pt r; r.y = v.x; r.x = v.y; return r;
}
__device__ __forceinline__ void st_stream_pt(pt* a, pt v)
{
    __stcs(reinterpret_cast<double2*>(a), make_double2(v.y, v.x));
}

// ---- TMA bulk copies (global -> shared) completing on an mbarrier: the state ring of k_advect_cert ----------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int cnt)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(cnt) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* b, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, unsigned long long* b)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, uint32_t parity)
{
    const uint32_t a = smem_u32(b);
    uint32_t done;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(a), "r"(parity) : "memory");
    } while (!done);
}
// the same with an L2 evict_first hint: the state stream is touch-once
__device__ __forceinline__ void tma_load_1d_stream(void* dst, const void* src, uint32_t bytes, unsigned long long* b)
{
    unsigned long long pol;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(b)), "l"(pol) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

// ---- tracking.py:44-49  _ccw_(A,B,C) = (Cy-Ay)*(Bx-Ax) > (By-Ay)*(Cx-Ax) ----
__device__ __forceinline__ bool ccw(pt A, pt B, pt C)
{
    return __dmul_rn(__dsub_rn(C.y, A.y), __dsub_rn(B.x, A.x)) >
           __dmul_rn(__dsub_rn(B.y, A.y), __dsub_rn(C.x, A.x));
}

// ---- tracking.py:51-58  intersect2Seg ---------------------------------------
// Branch-free: all four orientation tests are evaluated (no side effects, so
// Python's short-circuit `and` gives the same truth value).
__device__ __forceinline__ bool intersect2seg(pt A, pt B, pt C, pt D)
{
    return (ccw(A, C, D) != ccw(B, C, D)) & (ccw(A, B, C) != ccw(A, B, D));
}

// The same test when ccw(B,C,D) is already known.  In the U/V pick (si3_part_tracker.py:430-431) B, C, D
// are the cell's NE corner and two of its face points: that orientation is a property of the cell, computed
// once per grid by k_cell_bits with this very ccw() -- same operations, same rounding, same truth value.
__device__ __forceinline__ bool intersect2seg_pre(pt A, pt B, pt C, pt D, bool ccw_bcd)
{
    return (ccw(A, C, D) != ccw_bcd) & (ccw(A, B, C) != ccw(A, B, D));
}

// ---- locate.py:66-74  one edge (p1->p2) of the ray-casting parity test -------
__device__ __forceinline__ bool edge_toggles(double y, double x, pt p1, pt p2)
{
    const double ymin = (p2.y < p1.y) ? p2.y : p1.y;       // Python min(z1y,z2y)
    const double ymax = (p2.y > p1.y) ? p2.y : p1.y;
    const double xmax = (p2.x > p1.x) ? p2.x : p1.x;
    bool t = false;
    if (y > ymin && y <= ymax && x <= xmax) {
        // y>ymin && y<=ymax implies p1.y != p2.y, so xints is always refreshed here
        const double xints = __dadd_rn(
            __ddiv_rn(__dmul_rn(__dsub_rn(y, p1.y), __dsub_rn(p2.x, p1.x)), __dsub_rn(p2.y, p1.y)), p1.x);
        t = (p1.x == p2.x) || (x <= xints);
    }
    return t;
}

// ---- locate.py:49-78  IsInsideQuadrangle ------------------------------------
// The upstream loop runs 5 passes; pass 0 pairs vertex 0 with itself and can
// never toggle, leaving the four edges q0q1, q1q2, q2q3, q3q0.
__device__ __forceinline__ bool inside_quad(double y, double x, pt q0, pt q1, pt q2, pt q3)
{
    return edge_toggles(y, x, q0, q1) ^ edge_toggles(y, x, q1, q2) ^
           edge_toggles(y, x, q2, q3) ^ edge_toggles(y, x, q3, q0);
}

// ---- tracking.py:182-200  CrossedEdge -> 1 bottom, 2 right, 3 top, 4 left ----
// (4 is also the fall-through answer when no edge is met.)
__device__ __forceinline__ int crossed_edge(pt P1, pt P2, pt bl, pt br, pt ur, pt ul)
{
    if (intersect2seg(P1, P2, bl, br)) return 1;
    if (intersect2seg(P1, P2, br, ur)) return 2;
    if (intersect2seg(P1, P2, ur, ul)) return 3;
    return 4;
}

// ---- tracking.py:253-305  UpdtInd4NewCell as (dj,di) ---------------------------
__device__ __forceinline__ void cell_shift(int knhc, int& jT, int& iT)
{
    // 1 down, 2 right, 3 up, 4 left, 5 down-left, 6 down-right, 7 up-right, 8 up-left
    const int dj = (int)(knhc == 3 || knhc == 7 || knhc == 8) - (int)(knhc == 1 || knhc == 5 || knhc == 6);
    const int di = (int)(knhc == 2 || knhc == 6 || knhc == 7) - (int)(knhc == 4 || knhc == 5 || knhc == 8);
    jT += dj; iT += di;
}

// ---- tracking.py:62-93  Survive (first failing test wins; any >0 kills) -------
// tmask i1, ic = the CURRENT record's siconc (f4 widened exactly to f8, as the
// reference's f8 work array holds it, si3_part_tracker.py:372).
__device__ __forceinline__ bool killed(int jT, int iT, int Nj, int Ni,
                                       const int8_t* __restrict__ tmask, const float* __restrict__ ic,
                                       double rmin_conc)
{
    if (jT <= 1 || jT >= Nj - 2 || iT <= 1 || iT >= Ni - 2) return true;
    const int c = jT * Ni + iT;
    ST_CHECK_CELL(c, Ni + 1, Nj, Ni);
    const int zmt = __ldg(tmask + c) + __ldg(tmask + c + 1) + __ldg(tmask + c + Ni) +
                    __ldg(tmask + c - 1) + __ldg(tmask + c - Ni - 1);          // sic: [jT-1,iT-1]
    if (zmt < 5) return true;
    double s = __dadd_rn((double)__ldg(ic + c), (double)__ldg(ic + c + 1));
    s = __dadd_rn(s, (double)__ldg(ic + c + Ni));
    s = __dadd_rn(s, (double)__ldg(ic + c - 1));
    s = __dadd_rn(s, (double)__ldg(ic + c - Ni - 1));
    return __dmul_rn(0.2, s) < rmin_conc;
}

// ---- polar stereographic inverse (replaces util.py:413-429 -> cartopy/PROJ) ---
// PROJ iterates phi = pi/2 - 2 atan(t ((1-e sin phi)/(1+e sin phi))^(e/2)) to
// 1e-10 rad; here the same conformal->geodetic map is the 6-term series in the
// third flattening n (|truncation| < 1e-17 rad for WGS84), evaluated by Clenshaw
// with sin/cos of 2*chi obtained algebraically from t: one atan and one atan2
// per point instead of ~25 transcendentals.  Not bit-comparable with PROJ by
// construction (PROJ itself stops at 1e-10); tolerance in tests: 1e-9 degrees.
#define ST_LAT_DEG 8              // degree of the latitude polynomial: |error| < 2e-13 rad = 1e-11 degrees
#define ST_ANG_STEPS 64           // angle table: alpha_j = asin(j / ST_ANG_STEPS), j = 0 .. ST_ANG_LAST
#define ST_ANG_LAST 46            // > 64 / sqrt(2)
struct ProjConst {
    double k_t;        // 1000 / (a * akm1): km radius -> t = tan(pi/4 - chi/2)
    double c[6];       // series coefficients of sin(2k chi), k = 1..6
    double lon0_rad;   // central longitude
    double fill_lat, fill_lon;   // inv_stere of (-9999,-9999) km: what rows of idle buoys hold (:493)
    // inv_stere_fast works in degrees from the start:
    double w_scale;              // 8 k_t^2: w = 8 t^2 - 1 = r^2 w_scale - 1
    double lat_poly_deg[ST_LAT_DEG + 1];   // lat = 90 + t * sum_k lat_poly_deg[k] w^k for t <= 1/2 (lat >~ 37N)
    double oct_off[8], oct_sg[8];   // lon = oct_off[o] + oct_sg[o] * theta, o = swap | (cos<0)<<1 | (sin<0)<<2,
                                    // theta in [0, pi/4] the octant-folded angle; central longitude included
    int wrap_up;                 // lon0 > 0: only lon > 180 can occur; else only lon < -180
};

__device__ __forceinline__ pt inv_stere(pt yx, const ProjConst& pc)
{
    const double HALFPI = 1.5707963267948966, PI = 3.141592653589793, R2D = 57.29577951308232;
    const double t = sqrt(fma(yx.x, yx.x, yx.y * yx.y)) * pc.k_t;
    const double t2 = t * t;
    const double inv = 1.0 / (1.0 + t2);
    const double s = (1.0 - t2) * inv;              // sin(chi)
    const double c = 2.0 * t * inv;                 // cos(chi)
    const double s2 = 2.0 * s * c;                  // sin(2 chi)
    const double c2x2 = 2.0 * fma(-2.0 * s, s, 1.0);  // 2 cos(2 chi)
    double b2 = 0.0, b1 = pc.c[5];
#pragma unroll
    for (int k = 4; k >= 0; --k) { const double b0 = fma(c2x2, b1, pc.c[k] - b2); b2 = b1; b1 = b0; }
    const double phi = (HALFPI - 2.0 * atan(t)) + b1 * s2;
    double lam = (yx.x == 0.0 && yx.y == 0.0) ? 0.0 : atan2(yx.x, -yx.y);
    lam += pc.lon0_rad;
    if (lam > PI) lam -= 2.0 * PI;
    if (lam < -PI) lam += 2.0 * PI;
    pt r; r.y = phi * R2D; r.x = lam * R2D; return r;     // [lat, lon] degrees
}

// ---- fast inverse for the fused step (same map as inv_stere, fewer FP64-pipe slots) ------
// angle of a unit vector from a 47-entry table (1.5 KB: a 183-entry table with a shorter series put 8 % of the
// kernel's stall samples on the lookup): the octant-folded sine m = min(|s|,|c|) selects alpha_j = asin(j/64);
// the remainder delta (|delta| < 0.011) comes from sin(delta) = m cos(alpha_j) - M sin(alpha_j) and three terms
// of the asin series (error < 1e-15 rad).
struct __align__(16) AngEntry { double alpha, ca, sa, pad; };

// One Newton step on the MUFU seeds (~2^-20): relative error ~1e-12.  Enough here: both angle
// evaluations are direction-only (a common scale error of sin and cos cancels), and the error of
// t = rho * k_t moves the latitude by 2 t eps <= 1e-12 rad = 6e-11 degrees.
__device__ __forceinline__ double rcp_nr(double d)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double e = fma(-d, y, 1.0);
    return fma(y, e, y);
}
__device__ __forceinline__ double rsqrt_nr(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-(x * y), y, 1.0);
    return fma(0.5 * y, e, y);
}
// octant-folded angle theta in [0, pi/4] of the direction (s, c) (any common positive scale) and its
// octant o = swap | (c<0)<<1 | (s<0)<<2: the direction's angle from the +c axis towards +s is
//   swap=0,c>=0: +-theta   swap=1,c>=0: +-(pi/2 - theta)   swap=0,c<0: +-(pi - theta)   swap=1,c<0: +-(pi/2 + theta)
// with the sign of s.
// SM: `tab` is a shared-memory copy of the table (k_advect_cert: the lookup then never queues behind the global loads)
template <bool SM = false>
__device__ __forceinline__ double folded_angle(double s, double c, const AngEntry* __restrict__ tab, int& oct)
{
    const bool swp = fabs(s) > fabs(c);
    const double ms = swp ? c : s, Ms = swp ? s : c;           // signed; magnitudes taken at the uses
    int j = __double2int_rn(fabs(ms) * (double)ST_ANG_STEPS);
    j = min(max(j, 0), ST_ANG_LAST);
    double2 e0; double sa;                                                        // alpha, cos; sin
    if (SM) {
        e0 = *reinterpret_cast<const double2*>(tab + j);
        sa = tab[j].sa;
    } else {
        e0 = __ldg(reinterpret_cast<const double2*>(tab + j));
        sa = __ldg(reinterpret_cast<const double*>(tab + j) + 2);
    }
    const double sd = fma(fabs(ms), e0.y, -(fabs(Ms) * sa));                      // sin(delta), |delta| < 0.011
    const double z = sd * sd;
    const double d = fma(sd * z, fma(z, 3. / 40., 1. / 6.), sd);                  // asin: next term 15/336 sd^7 < 1e-15 rad
    oct = (int)swp | ((int)(c < 0.0) << 1) | ((__double2hiint(s) >> 31) & 1) << 2;
    return e0.x + d;
}
template <bool SM = false>
__device__ __forceinline__ double angle_of_unit(double s, double c, const AngEntry* __restrict__ tab)
{
    const double HALFPI = 1.5707963267948966, PI = 3.141592653589793;
    int oct;
    double th = folded_angle<SM>(s, c, tab, oct);
    if (oct & 1) th = HALFPI - th;
    if (oct & 2) th = PI - th;
    return copysign(th, s);
}

template <bool SM = false>
__device__ __forceinline__ pt inv_stere_fast(pt yx, const ProjConst& pc, const AngEntry* __restrict__ tab)
{
    const double R2D = 57.29577951308232;
    const double r2 = fma(yx.x, yx.x, yx.y * yx.y);
    const double rinv = (r2 > 0.0) ? rsqrt_nr(r2) : 0.0;          // pole: lam = 0 like PROJ
    const double t = (r2 * rinv) * pc.k_t;
    const double w = fma(r2, pc.w_scale, -1.0);                    // 8 t^2 - 1
    pt r;
    if (w <= 1.0) {
        // latitudes above ~37N: phi - pi/2 is odd in t, phi = pi/2 + t Q(8 t^2 - 1) with Q a degree-8
        // polynomial fitted at st_create for the ellipsoid (|error| < 2e-13 rad = 1e-11 degrees, see make_proj)
        double q = pc.lat_poly_deg[ST_LAT_DEG];
#pragma unroll
        for (int k = ST_LAT_DEG - 1; k >= 0; --k) q = fma(q, w, pc.lat_poly_deg[k]);
        r.y = fma(t, q, 90.0);
    } else {
        const double t2 = t * t;
        const double inv = rcp_nr(1.0 + t2);
        const double s = (1.0 - t2) * inv;              // sin(chi)
        const double c = (t + t) * inv;                 // cos(chi)
        const double chi = angle_of_unit<SM>(s, c, tab);
        const double s2 = (s + s) * c;                  // sin(2 chi)
        const double c2x2 = fma(-4.0 * s, s, 2.0);      // 2 cos(2 chi)
        double b2 = 0.0, b1 = pc.c[4];                  // c[5] = 6e-16 rad is dropped
#pragma unroll
        for (int k = 3; k >= 0; --k) { const double b0 = fma(c2x2, b1, pc.c[k] - b2); b2 = b1; b1 = b0; }
        r.y = fma(b1, s2, chi) * R2D;
    }
    int oct;
    const double th = folded_angle<SM>(yx.x * rinv, -(yx.y * rinv), tab, oct);
    double lon = fma(th, pc.oct_sg[oct], pc.oct_off[oct]);
    if (pc.wrap_up) { if (lon > 180.0) lon -= 360.0; }
    else            { if (lon < -180.0) lon += 360.0; }
    r.x = lon;
    return r;
}

// x / 1000 correctly rounded without the division sequence: q = RN(a r), r = RN(1/1000),
// then one exact-residual correction (Markstein).  Identical to __ddiv_rn(a, 1000.) for every
// finite a in the range velocities can produce (checked against IEEE division on 3e8 random
// and 1.5e9 near-midpoint operands, and on the device in tests/test_gpu_parity.py);
// differs only for a = +-inf (NaN instead of inf).
__device__ __forceinline__ double div1000(double a)
{
    const double r = 1.0 / 1000.0;
    const double q = __dmul_rn(a, r);
    const double e = __fma_rn(-1000.0, q, a);
    return __fma_rn(e, r, q);
}

// forward, for the grid-preparation helpers (ncio.py:50-53,86-89)
struct ProjFwdConst { double a_akm1_km; double e; double lon0_rad; };
__device__ __forceinline__ pt fwd_stere(pt latlon, const ProjFwdConst& pc)
{
    const double D2R = 0.017453292519943295, HALFPI = 1.5707963267948966;
    const double phi = latlon.y * D2R;
    const double lam = latlon.x * D2R - pc.lon0_rad;
    double sphi, cphi; sincos(phi, &sphi, &cphi);
    const double es = pc.e * sphi;
    // t = tan(pi/4 - phi/2) / ((1-es)/(1+es))^(e/2); tan(pi/4-phi/2) = cos(phi)/(1+sin(phi))
    const double t = (fabs(phi - HALFPI) < 1e-15) ? 0.0
                   : (cphi / (1.0 + sphi)) * exp(pc.e * atanh(es));   // ((1+es)/(1-es))^(e/2) = exp(e atanh(es))
    const double rho = pc.a_akm1_km * t;
    double sl, cl; sincos(lam, &sl, &cl);
    pt r; r.y = -rho * cl; r.x = rho * sl; return r;
}

// ---- util.py:85-103  Haversine distance [km], same expression order ------------
__device__ __forceinline__ double haversine_km(double plat, double plon, double xlat, double xlon)
{
    const double to_rad = 3.141592653589793 / 180.;
    const double a1 = sin(__dmul_rn(0.5, __dmul_rn(__dsub_rn(xlat, plat), to_rad)));
    const double a2 = sin(__dmul_rn(0.5, __dmul_rn(__dsub_rn(xlon, plon), to_rad)));
    const double a3 = __dmul_rn(cos(__dmul_rn(xlat, to_rad)), cos(__dmul_rn(plat, to_rad)));
    const double h = __dadd_rn(__dmul_rn(a1, a1), __dmul_rn(__dmul_rn(a3, a2), a2));
    return __dmul_rn(2. * 6360., asin(sqrt(h)));
}

}  // namespace st
