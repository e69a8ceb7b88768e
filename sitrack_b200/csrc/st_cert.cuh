// st_cert.cuh -- k_advect_cert: the step with a CERTIFIED fast path (default kernel, variant 0).
//
// k_advect_warp gathers eight f8 points (128 B) per buoy to run the reference's two U/V-pick segment tests
// and an orientation filter for the inside test -- and is bound by the latency and the L1 wavefronts of those
// gathers, not by HBM.  Yet ~80 % of the buoys sit well inside their cell, well away from the two lines that
// separate "nearest U is the west one" / "nearest V is the south one", and stay inside after the step.  For
// those the reference's decisions can be PROVEN from far less data:
//
//   * k_cell_frames (once per grid) gives every host cell an affine frame  (s,t) = M (P - O)  in which the
//     cell is ~[-1/2,1/2]^2: origin O and the 2x2 matrix M as f32 (32 B, two LDG.128), plus two margins
//     packed as bf16 (4 B): `msep` and `hin`.  The frame is DEFINED by those f32 numbers; the kernel that builds
//     it then measures, in f8, where the cell's eight real points (4 corners, 2 V-points, 2 U-points) lie in
//     that frame and derives the margins from them, verifying every claim below on the corners of the
//     certified regions (each claim is the sign of a function affine in P, so its extreme over a rectangle
//     is at a corner).
//   * the step evaluates (s,t) of P and (s',t') of the new position in f32 (6 FFMA) and certifies
//       pick :  |s|,|t| < hin  and  |s|,|t| > msep   =>  llum1 = (s < 0), llvm1 = (t < 0)   (si3_part_tracker.py:430-441)
//       stay :  |s'|,|t'| < hin                      =>  IsInsideQuadrangle(new position) is True   (locate.py:49-78)
//     The new position itself is always the reference's f8 Euler arithmetic on the picked f4 velocities.
//   * a lane whose pick is not certified is queued (index only) and the dense X pass runs the reference's own
//     step on it (exact pick, exact inside test, walk); a lane whose pick is certified but whose stay is not
//     (it crossed, or it is within the margin of an edge) is queued with P and Pn and the dense W pass runs the
//     reference's inside test and, if outside, the walk -- exactly k_advect_warp's dense pass.
//
// Why a certified decision equals the reference's (proof obligations, all checked per cell by k_cell_frames):
//   (0) M has positive determinant, so every orientation test ccw(A,B,C) = [orient(A,B,C) > 0] has the same exact
//       sign in frame coordinates, and orient_frame = orient_km * det M.  The reference evaluates orient_km in f8
//       from exact inputs with at most 6 2^-53 D^2 of error (D = largest coordinate difference involved); the cell
//       is admitted only if D^2 det M <= 2^10, i.e. the f8 sign is the exact one whenever |orient_frame| >= 2^-40.
//   (1) f32 evaluation: with kappa = max row-wise sum |M||M^-1| (1 for an axis-aligned square, < 2.5 for sane
//       cells, admitted up to 64) the computed frame coordinates of P and of the new position are within
//       eps = 2^-20 kappa + 2^-34 sum|M_ij| of the exact ones whenever they pass the `< hin <= 1/2` tests (the bound
//       scales with the true magnitudes, so coordinates that are truly large cannot come out small).
//   (2) pick, U side: llum1 = intersect2Seg(P, UR, v0, v1) = [ccw(P,v0,v1) != ccw(UR,v0,v1)] and
//       [ccw(P,UR,v0) != ccw(P,UR,v1)].  The cell bit ccw(UR,v0,v1) (evaluated with the reference's own expression)
//       must be False.  The line v0->v1 stays within |s| <= sU for |t| <= 0.55; msep >= sU + 2 eps + 2^-28, so a
//       certified s > msep puts P on UR's side (first bracket False => llum1 False), and a certified s < -msep
//       puts it on the other side with orient >= 2^-29; the second bracket is then verified True on the four
//       corners of {-h' <= s <= -msep + eps, |t| <= h'} (h' = hin + eps): orient(P,UR,v1) >= 2^-40 and
//       orient(P,UR,v0) <= -2^-40.  V side alike with u0,u1 (bit ccw(UR,u0,u1) must be True).
//   (3) stay: on the four corners of [-h',h']^2 the point is to the left of the four directed edges
//       BL->BR->UR->UL->BL by at least 2^-17 km^2 of exact cross product; with the cell convex, anticlockwise, edges
//       < 2^10 km and coordinates < 2^17 km (k_cell_bits bit 2, filter_ok) that is the premise under which
//       inside_margin's argument shows the reference's ray casting answers True (st_advect.cu).
// A cell that fails any check gets hin = -1: nothing is ever certified in it and every buoy there takes the
// exact path.  NaN/Inf velocities make s',t' NaN/Inf: every comparison is written so that they fail to certify.
//
// tests: tests/test_gpu_parity.py::test_cert_* (certified decisions against the exact ones on millions of points
// placed at the margins; fraction certified) and every bit-exact trajectory test, which now runs through here.
#pragma once
#include "st_kernels.h"

namespace st {

// the certified part of the step for one lane: frame coordinates, pick, Euler step, stay test
struct CertLane { bool pick_ok, in_ok, left, below; };

template <int UV>
__device__ __forceinline__ CertLane cert_eval(const float4 f0, const float4 f1, pt P,
                                              float uL, float uR, float vB, float vT, double rdt, float kdt,
                                              pt& Pn)
{
    // f0 = {oy, ox, a, b}, f1 = {c, d, bf16 hin | bf16 msep, bf16 es | bf16 et}:  s = a dx + b dy + es,  t = c dx + d dy + et
    const unsigned mw = __float_as_uint(f1.z), ew = __float_as_uint(f1.w);
    const float hin = __uint_as_float(mw & 0xffff0000u), msep = __uint_as_float(mw << 16);
    const float es = __uint_as_float(ew & 0xffff0000u), et = __uint_as_float(ew << 16);
    const float dy = __double2float_rn(__dsub_rn(P.y, (double)f0.x));
    const float dx = __double2float_rn(__dsub_rn(P.x, (double)f0.y));
    const float s = __fmaf_rn(f0.z, dx, __fmaf_rn(f0.w, dy, es));
    const float t = __fmaf_rn(f1.x, dx, __fmaf_rn(f1.y, dy, et));
    const float as = fabsf(s), at = fabsf(t);
    CertLane r;
    r.left = s < 0.0f; r.below = t < 0.0f;
    double zU, zV; float zUf, zVf;
    if (UV == 1) {
        r.pick_ok = (as < hin) & (at < hin) & (as > msep) & (at > msep);
        zUf = r.left ? uL : uR;                      // si3_part_tracker.py:432-435 when certified
        zVf = r.below ? vB : vT;                     // :436-439
        zU = (double)zUf; zV = (double)zVf;
    } else {
        r.pick_ok = true;                            // :423-425 involves no decision
        zU = __dmul_rn(0.5, __dadd_rn((double)uR, (double)uL));
        zV = __dmul_rn(0.5, __dadd_rn((double)vT, (double)vB));
        zUf = __double2float_rn(zU); zVf = __double2float_rn(zV);
    }
    Pn.x = __dadd_rn(P.x, div1000(__dmul_rn(zU, rdt)));          // :452-458, the reference's f8 arithmetic
    Pn.y = __dadd_rn(P.y, div1000(__dmul_rn(zV, rdt)));
    const float ddx = __fmul_rn(zUf, kdt), ddy = __fmul_rn(zVf, kdt);
    const float s2 = __fmaf_rn(f0.z, ddx, __fmaf_rn(f0.w, ddy, s));
    const float t2 = __fmaf_rn(f1.x, ddx, __fmaf_rn(f1.y, ddy, t));
    r.in_ok = (fabsf(s2) < hin) & (fabsf(t2) < hin);
    if (UV != 1) r.in_ok = r.in_ok & (as < hin) & (at < hin);    // bound (1) needs P itself in range
    return r;
}

// Launch shape and data flow.  The memory-system skeleton of the step (tools/micro/stream_skeleton.cu: the same five
// streams, no gathers, no arithmetic) runs at 7.1 TB/s as a plain one-thread-per-buoy grid at full occupancy and at
// 5.1-6.5 TB/s as persistent one-warp CTAs with 16-32 warps per SM: what this access mix needs is bytes in flight,
// i.e. resident threads.  So the step is two plain kernels:
//   k_advect_cert  one thread per buoy, nothing but the certified fast path (48 registers, no shared memory): state
//                  in, frame + four face velocities gathered, pick / Euler step / stay decided, rows and state out.
//                  A lane it cannot certify is only MARKED: one ballot word per tile of 32 buoys for "stay not
//                  certified" (W), one for "pick not certified" (X); a W lane's old position goes to a scratch array
//                  at its own index (its state already holds the new one), an X lane's state is left untouched.
//   k_walk         one thread per marked lane, dense: a block takes ST_WALK_TILES tiles, scans their ballot words and
//                  deals the marked lanes out to its threads.  W: the reference's inside test on the new position and,
//                  when outside, CrossedEdge / NewHostCell / UpdtInd4NewCell / Survive.  X: the reference's whole
//                  step (exact_lane).
// `alive` is bit 31 of the cell's jT word.
constexpr int ST_CERT_BLOCK = 256;
#ifndef ST_CERT_MINBLK
#define ST_CERT_MINBLK 8                                          // full occupancy: 2048 resident threads per SM at 32 registers (12 B of spill; 6 blocks x 40 registers measured slower)
#endif
#ifndef ST_WALK_TILES
#define ST_WALK_TILES 24
#endif
constexpr int ST_WALK_BLOCK = 128;

__device__ __forceinline__ float4 ldg_f4_keep(const float4* a)
{
    float4 r;
    asm("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
        : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(a), "l"(l2_keep_policy()));
    return r;
}

// X lanes of k_walk: the reference's whole step with its own tests.  k_advect_cert has already stored a TENTATIVE new
// position, rows included, from the frame's uncertified pick (whole sectors, like every other lane); it is replaced
// only when the reference's pick gives a different one -- a lone 16-byte store is expensive (see k_advect_cert).
template <int UV, int ROWS>
__device__ __forceinline__ void exact_lane(const AdvectGrid& g, const float* __restrict__ u, const float* __restrict__ v,
                                           const float* __restrict__ ic, const BuoyState& s, const StepOut& o,
                                           const WalkScratch& q, unsigned p)
{
    const pt P = ld_stream_pt(q.P + p), Pt = ld_stream_pt(s.pos + p);
    int2 cc = __ldcs(s.cell + p);
    const int j0 = cc.x, i0 = cc.y;
    const int Ni = g.Ni;
    const int c = cc.x * Ni + cc.y;
    ST_CHECK_CELL(c, Ni + 1, g.Nj, Ni);
    const pt bl = ldg_pt(g.F, c - Ni - 1), br = ldg_pt(g.F, c - Ni);
    const pt ul = ldg_pt(g.F, c - 1),      ur = ldg_pt(g.F, c);
    const pt v0 = ldg_pt(g.V, c - Ni), v1 = ldg_pt(g.V, c);
    const pt u0 = ldg_pt(g.U, c - 1),  u1 = ldg_pt(g.U, c);
    const float uL = __ldg(u + c - 1), uR = __ldg(u + c);
    const float vB = __ldg(v + c - Ni), vT = __ldg(v + c);
    const bool llum1 = intersect2seg(P, ur, v0, v1);               // si3_part_tracker.py:430
    const bool llvm1 = intersect2seg(P, ur, u0, u1);               // :431
    const double zU = (double)(llum1 ? uL : uR);
    const double zV = (double)(llvm1 ? vB : vT);
    pt Pn;
    Pn.x = __dadd_rn(P.x, div1000(__dmul_rn(zU, g.rdt)));
    Pn.y = __dadd_rn(P.y, div1000(__dmul_rn(zV, g.rdt)));
    if (__double_as_longlong(Pn.x) != __double_as_longlong(Pt.x) || __double_as_longlong(Pn.y) != __double_as_longlong(Pt.y)) {
        st_stream_pt(s.pos + p, Pn);
        if (o.yx) { if (ROWS == 0) st_stream_pt(o.yx + p, Pn); else put_row_yx(o, p, Pn); }
        if (o.latlon) {
            const pt ll = inv_stere_fast(Pn, g.proj, g.atab);
            if (ROWS == 0) st_stream_pt(o.latlon + p, ll); else put_row_pt(o.latlon, p, ll, o.f4);
        }
    }
    if (!inside_quad_flat(Pn.y, Pn.x, bl, br, ur, ul)) {
        int8_t a2 = 1;
        walk_cell(g, ic, P, Pn, cc.x, cc.y, a2);
        if (!a2) cc.x |= ST_DEAD_BIT;
        if (cc.x != j0 || cc.y != i0) __stcs(s.cell + p, cc);
        if (!a2) s.alive[p] = 0;
    }
}

template <int UV, bool WIN, int ROWS>
__global__ void __launch_bounds__(ST_CERT_BLOCK, ST_CERT_MINBLK)
k_advect_cert(const AdvectGrid g, const float* __restrict__ u, const float* __restrict__ v,
              BuoyState s, int jrec, StepOut o, WalkScratch q)
{
    const long long p = (long long)blockIdx.x * ST_CERT_BLOCK + threadIdx.x;
    const bool valid = p < s.nP;
    const int Ni = g.Ni;
    pt P = {ST_FILL, ST_FILL};
    int2 c2 = make_int2(ST_DEAD_BIT | 2, 2);
    if (valid) { P = ld_stream_pt(s.pos + p); c2 = __ldcs(s.cell + p); }
    const bool al = c2.x >= 0;
    bool active = al;
    bool prestart = false;
    if (WIN && active) {
        const int f = s.rec_first[p], l = s.rec_last[p];
        prestart = (jrec + 1 == f);
        active = (jrec >= f) && (jrec <= l);
    }
    const int c = active ? c2.x * Ni + c2.y : 2 * Ni + 2;
    ST_CHECK_CELL(c, Ni + 1, g.Nj, Ni);
    const float4* fr = g.frames + 2 * (size_t)c;
    const float4 f0 = ldg_f4_keep(fr), f1 = ldg_f4_keep(fr + 1);
    const float uL = __ldg(u + c - 1), uR = __ldg(u + c);
    const float vB = __ldg(v + c - Ni), vT = __ldg(v + c);
    const float kdt = __double2float_rn(g.rdt / 1000.0);
    pt Pn;
    const CertLane cl = cert_eval<UV>(f0, f1, P, uL, uR, vB, vT, g.rdt, kdt, Pn);
    const bool goX = active && !cl.pick_ok;                       // k_walk: Pn is tentative, the reference's own step decides
    const bool goW = active && cl.pick_ok && !cl.in_ok;           // k_walk: Pn is final; inside test + walk by the reference's tests
    pt outp = {ST_FILL, ST_FILL};
    int8_t m = 0;
    if (active) {
        outp = Pn; m = 1;
        st_stream_pt(s.pos + p, outp);
    } else if (WIN && prestart) {
        outp = P; m = 1;
    }
    // The old position of a marked lane goes to the scratch array in WHOLE 32-byte sectors: a lane and its pair neighbour
    // both store when either is marked.  A lone 16-byte store leaves a half-written sector behind, and those cost this
    // step 80 us in k_advect_cert and 95 us in k_walk (read-modify-write at eviction; measured, profiles/README.md).
    { const int mk = (int)(goW | goX); if (__shfl_xor_sync(0xffffffffu, mk, 1) | mk) { if (valid) q.P[p] = P; } }
    if (valid) {
        if (o.mask) __stcs(o.mask + p, m);
        if (ROWS == 0) { if (o.yx) st_stream_pt(o.yx + p, outp); }
        else           { if (o.yx) put_row_yx(o, p, outp); }
        if (o.latlon) {
            pt ll; ll.y = g.proj.fill_lat; ll.x = g.proj.fill_lon;                    // :493 on a fill row
            if (m) ll = inv_stere_fast(outp, g.proj, g.atab);
            if (ROWS == 0) st_stream_pt(o.latlon + p, ll);
            else           put_row_pt(o.latlon, p, ll, o.f4);
        }
    }
    const unsigned balW = __ballot_sync(0xffffffffu, goW);
    const unsigned balX = __ballot_sync(0xffffffffu, goX);
    const unsigned balA = __ballot_sync(0xffffffffu, valid && al);
    if ((threadIdx.x & 31) == 0) { q.maskW[p >> 5] = balW; q.maskX[p >> 5] = balX; q.maskA[p >> 5] = balA; }
}

template <int UV, int ROWS>
__global__ void __launch_bounds__(ST_WALK_BLOCK)
k_walk(const AdvectGrid g, const float* __restrict__ u, const float* __restrict__ v, const float* __restrict__ ic,
       BuoyState s, StepOut o, WalkScratch q, int ntiles)
{
    constexpr int T = ST_WALK_TILES;
    static_assert(T <= 32, "one warp scans the ballot words of a block");
    __shared__ unsigned sMask[2][T];
    __shared__ int sPre[2][T + 1];                                // exclusive prefix of the counts: W, then X
    const int tile0 = (int)blockIdx.x * T;
    if (threadIdx.x < 32) {
        const int t = tile0 + (int)threadIdx.x;
        const bool in = (int)threadIdx.x < T && t < ntiles;
        const unsigned mw = in ? q.maskW[t] : 0u, mx = (in && UV == 1) ? q.maskX[t] : 0u;
        if (o.n_alive) {                                          // buoys that were alive when the step began
            const int na = __reduce_add_sync(0xffffffffu, in ? __popc(q.maskA[t]) : 0);
            if (threadIdx.x == 0 && na) atomicAdd(o.n_alive, (unsigned long long)na);
        }
        int cw = __popc(mw), cx = __popc(mx);
        int iw = cw, ix = cx;                                     // inclusive warp scans
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int aw = __shfl_up_sync(0xffffffffu, iw, d), ax = __shfl_up_sync(0xffffffffu, ix, d);
            if ((int)threadIdx.x >= d) { iw += aw; ix += ax; }
        }
        if (threadIdx.x < T) {
            sMask[0][threadIdx.x] = mw; sMask[1][threadIdx.x] = mx;
            sPre[0][threadIdx.x + 1] = iw; sPre[1][threadIdx.x + 1] = ix;
        }
        if (threadIdx.x == 0) { sPre[0][0] = 0; sPre[1][0] = 0; }
    }
    __syncthreads();
    const int nW = sPre[0][T], nX = sPre[1][T];
    const int Ni = g.Ni;
    for (int e = threadIdx.x; e < nW + nX; e += ST_WALK_BLOCK) {
        const int kind = e >= nW;                                 // 0: W entry, 1: X entry
        const int r = kind ? e - nW : e;
        int lo = 0;                                               // tile whose prefix range holds r
#pragma unroll
        for (int step = 16; step > 0; step >>= 1)
            if (lo + step < T && sPre[kind][lo + step] <= r) lo += step;
        const unsigned lane_of = __fns(sMask[kind][lo], 0, r - sPre[kind][lo] + 1);
        const unsigned p = (unsigned)(tile0 + lo) * 32u + lane_of;
        if (kind == 0) {
            const pt A = ld_stream_pt(q.P + p), B = ld_stream_pt(s.pos + p);
            int2 cc = __ldcs(s.cell + p);
            const int j0 = cc.x, i0 = cc.y;
            const int c = cc.x * Ni + cc.y;
            if (!inside_quad_flat(B.y, B.x, ldg_pt(g.F, c - Ni - 1), ldg_pt(g.F, c - Ni), ldg_pt(g.F, c),
                                  ldg_pt(g.F, c - 1))) {
                int8_t a2 = 1;
                walk_cell(g, ic, A, B, cc.x, cc.y, a2);
                if (!a2) cc.x |= ST_DEAD_BIT;
                if (cc.x != j0 || cc.y != i0) __stcs(s.cell + p, cc);
                if (!a2) s.alive[p] = 0;
            }
        } else {
            if (UV == 1) exact_lane<UV, ROWS>(g, u, v, ic, s, o, q, p);
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// k_cell_frames: the frame and the margins of every host cell, once per grid (st_create).  FP64 throughout;
// see the head of this file for what is verified.  Output per cell c = jT*Ni + iT:
//   frames[2c]   = {oy, ox, a, b}     frames[2c+1] = {c, d, mw, ew}     (f32; s = a dx + b dy + es, t = c dx + d dy + et)
//   mw = bf16(hin) << 16 | bf16(msep)   (hin rounded down, msep rounded up; hin = -1: never certify)
//   ew = bf16(es) << 16 | bf16(et)      (the offsets that centre the frame, truncated to bf16: 32 B per cell, one sector)
// stats[0] = cells admitted, stats[1] = cells examined.
// ---------------------------------------------------------------------------------------------------
struct FramePt { double s, t; };
__device__ __forceinline__ double orient2(FramePt a, FramePt b, FramePt c)
{
    return (b.s - a.s) * (c.t - a.t) - (b.t - a.t) * (c.s - a.s);
}
__device__ __forceinline__ unsigned bf16_down(float x)           // largest bf16 <= x (x >= 0)
{
    return __float_as_uint(x) >> 16;
}
__device__ __forceinline__ unsigned bf16_up(float x)             // smallest bf16 >= x (x >= 0)
{
    const unsigned b = __float_as_uint(x);
    return (b >> 16) + ((b & 0xffffu) ? 1u : 0u);
}

__global__ void __launch_bounds__(ST_BLOCK)
k_cell_frames(const AdvectGrid g, float4* __restrict__ frames, unsigned long long* __restrict__ stats)
{
    const long long c = (long long)blockIdx.x * ST_BLOCK + threadIdx.x;
    const long long n = (long long)g.Nj * g.Ni;
    if (c >= n) return;
    const int Ni = g.Ni;
    const int j = (int)(c / Ni), i = (int)(c % Ni);
    const unsigned NEVER = 0xbf800000u;                            // hin = -1 (bf16 of -1 is exact), msep = 0
    float4 o0 = make_float4(0.f, 0.f, 0.f, 0.f), o1 = o0;
    unsigned mw = NEVER;
    const bool uv = (g.U != nullptr) && (g.V != nullptr) && g.uv_strategy == 1;
    if (j >= 1 && i >= 1 && (g.cellbits[c] & 4) && g.filter_ok) {
        const pt bl = g.F[c - Ni - 1], br = g.F[c - Ni], ur = g.F[c], ul = g.F[c - 1];
        const double cy = 0.25 * ((bl.y + br.y) + (ur.y + ul.y)), cx = 0.25 * ((bl.x + br.x) + (ur.x + ul.x));
        const float oyf = (float)cy, oxf = (float)cx;
        const double oy = (double)oyf, ox = (double)oxf;
        // axes of the cell: es along +s (west -> east), et along +t (south -> north)
        const double esx = 0.5 * ((br.x + ur.x) - (bl.x + ul.x)), esy = 0.5 * ((br.y + ur.y) - (bl.y + ul.y));
        const double etx = 0.5 * ((ul.x + ur.x) - (bl.x + br.x)), ety = 0.5 * ((ul.y + ur.y) - (bl.y + br.y));
        const double det = esx * ety - esy * etx;
        bool ok = det > 0.0 && isfinite(det);
        const float af = (float)(ety / det), bf = (float)(-etx / det), cf = (float)(-esy / det), df = (float)(esx / det);
        const double a = af, b = bf, cc = cf, d = df;             // the frame IS these f32 numbers
        // (truncated to bf16: like a..d these numbers DEFINE the frame, everything below is measured in it)
        const float esf = __uint_as_float(__float_as_uint((float)(-(a * (cx - ox) + b * (cy - oy)))) & 0xffff0000u);
        const float etf = __uint_as_float(__float_as_uint((float)(-(cc * (cx - ox) + d * (cy - oy)))) & 0xffff0000u);
        const double es0 = esf, et0 = etf;
        const double detM = a * d - b * cc;
        ok = ok && detM > 0.0 && isfinite(detM);
        auto loc = [&](pt q) { FramePt r; const double dx = q.x - ox, dy = q.y - oy;
                               r.s = a * dx + b * dy + es0; r.t = cc * dx + d * dy + et0; return r; };
        const FramePt qbl = loc(bl), qbr = loc(br), qur = loc(ur), qul = loc(ul);
        // (1) evaluation error of the f32 frame coordinates
        const double kap = fmax(fabs(a * d) + fabs(b * cc) + 2.0 * fabs(a * b), fabs(a * d) + fabs(b * cc) + 2.0 * fabs(cc * d)) / detM;
        const double eps = 9.5367431640625e-07 * kap + 5.820766091346741e-11 * (fabs(a) + fabs(b) + fabs(cc) + fabs(d));
        ok = ok && kap <= 64.0 && eps < 1e-3;
        // (0) the reference's f8 orientation tests have the exact sign beyond 2^-40 in frame units
        double D = 0.0;
        { const pt q[4] = {bl, br, ur, ul};
          for (int k = 0; k < 4; ++k) for (int l = k + 1; l < 4; ++l) D = fmax(D, fmax(fabs(q[k].x - q[l].x), fabs(q[k].y - q[l].y))); }
        ok = ok && (4.0 * D * D * detM <= 1024.0);                 // points of the tests lie within the cell's hull (P certified inside)
        // deviation of the corners from the ideal square
        double dev = 0.0;
        dev = fmax(dev, fmax(fabs(qbl.s + 0.5), fabs(qbl.t + 0.5)));
        dev = fmax(dev, fmax(fabs(qbr.s - 0.5), fabs(qbr.t + 0.5)));
        dev = fmax(dev, fmax(fabs(qur.s - 0.5), fabs(qur.t - 0.5)));
        dev = fmax(dev, fmax(fabs(qul.s + 0.5), fabs(qul.t - 0.5)));
        double msep = 0.0;
        FramePt qv0 = {0, 0}, qv1 = {0, 0}, qu0 = {0, 0}, qu1 = {0, 0};
        if (uv) {
            qv0 = loc(g.V[c - Ni]); qv1 = loc(g.V[c]); qu0 = loc(g.U[c - 1]); qu1 = loc(g.U[c]);
            dev = fmax(dev, fmax(fabs(qv0.s), fabs(qv0.t + 0.5)));
            dev = fmax(dev, fmax(fabs(qv1.s), fabs(qv1.t - 0.5)));
            dev = fmax(dev, fmax(fabs(qu0.s + 0.5), fabs(qu0.t)));
            dev = fmax(dev, fmax(fabs(qu1.s - 0.5), fabs(qu1.t)));
            // (2) the cell bits of the reference's own ccw(UR, ., .)
            ok = ok && ((g.cellbits[c] & 3) == 2);
            const double hv = qv1.t - qv0.t, hu = qu1.s - qu0.s;
            ok = ok && hv >= 0.5 && hu >= 0.5;
            // extent of the separator lines over the certified range of the other coordinate
            const double sl = (qv1.s - qv0.s) / hv, tl = (qu1.t - qu0.t) / hu;
            const double sU = fmax(fabs(qv0.s + sl * (-0.55 - qv0.t)), fabs(qv0.s + sl * (0.55 - qv0.t)));
            const double tV = fmax(fabs(qu0.t + tl * (-0.55 - qu0.s)), fabs(qu0.t + tl * (0.55 - qu0.s)));
            msep = fmax(sU, tV) + 2.0 * eps + 3.725290298461914e-09;
        }
        ok = ok && dev < 0.05 && msep < 0.2 && isfinite(dev) && isfinite(msep);
        const float msep_f = __uint_as_float(bf16_up((float)msep) << 16);
        const double msep_r = (double)msep_f;                      // the margin the kernel will actually use
        // (2b), (3): search the shrink margin, verify on rectangle corners
        const double E40 = 9.094947017729282e-13;                  // 2^-40
        const double Ein = 7.62939453125e-06 * detM * 1.0000001 + E40;   // 2^-17 km^2 in frame units
        float hin_f = -1.0f;
        if (ok) {
            double m = 8.0 * dev + 4.0 * eps + 9.5367431640625e-07;
            for (int it = 0; it < 6 && hin_f < 0.0f; ++it, m *= 2.0) {
                if (m >= 0.25) break;
                const float hf = __uint_as_float(bf16_down((float)(0.5 - m)) << 16);
                const double h = (double)hf + eps;                 // true coordinates of a certified point lie within +-h
                bool good = h < 0.55;
                // (3) stay: left of the four directed edges by Ein at the four corners of [-h,h]^2
                const FramePt sq[4] = {{-h, -h}, {h, -h}, {h, h}, {-h, h}};
                for (int k = 0; k < 4 && good; ++k)
                    good = orient2(qbl, qbr, sq[k]) >= Ein && orient2(qbr, qur, sq[k]) >= Ein &&
                           orient2(qur, qul, sq[k]) >= Ein && orient2(qul, qbl, sq[k]) >= Ein;
                if (uv && good) {
                    const double sm = -msep_r + eps;               // certified s < -msep  =>  true s < sm
                    good = sm < 0.0 && -h < sm;
                    // (2) the segment P->UR separates v0 from v1 for P on the west side ...
                    const FramePt ru[4] = {{-h, -h}, {sm, -h}, {sm, h}, {-h, h}};
                    for (int k = 0; k < 4 && good; ++k)
                        good = orient2(ru[k], qur, qv1) >= E40 && orient2(ru[k], qur, qv0) <= -E40 &&
                               // ... and P is strictly on the far side of v0->v1 from UR (first bracket)
                               orient2(ru[k], qv0, qv1) >= E40;
                    // and u0 from u1 for P on the south side
                    const FramePt rv[4] = {{-h, -h}, {h, -h}, {h, sm}, {-h, sm}};
                    for (int k = 0; k < 4 && good; ++k)
                        good = orient2(rv[k], qur, qu1) <= -E40 && orient2(rv[k], qur, qu0) >= E40 &&
                               orient2(rv[k], qu0, qu1) <= -E40;
                    // east / north sides: P strictly on UR's side of the separator (first bracket False)
                    const FramePt re[4] = {{-sm, -h}, {h, -h}, {h, h}, {-sm, h}};
                    for (int k = 0; k < 4 && good; ++k) good = orient2(re[k], qv0, qv1) <= -E40;
                    const FramePt rn[4] = {{-h, -sm}, {h, -sm}, {h, h}, {-h, h}};
                    for (int k = 0; k < 4 && good; ++k) good = orient2(rn[k], qu0, qu1) >= E40;
                }
                if (good) hin_f = hf;
            }
        }
        if (hin_f > 0.0f) {
            mw = (__float_as_uint(hin_f) & 0xffff0000u) | (__float_as_uint(msep_f) >> 16);
            o0 = make_float4(oyf, oxf, af, bf);
            o1 = make_float4(cf, df, __uint_as_float(mw), __uint_as_float(__float_as_uint(esf) | (__float_as_uint(etf) >> 16)));
            atomicAdd(stats, 1ull);
        }
        atomicAdd(stats + 1, 1ull);
    }
    if (mw == NEVER) o1.z = __uint_as_float(NEVER);
    frames[2 * c] = o0; frames[2 * c + 1] = o1;
}

}  // namespace st

namespace st {
// k_cert_selftest: for n arbitrary (P, host cell, four face velocities) evaluate the certified decisions AND
// the reference's exact ones.  flags: bit0 pick certified, bit1 stay certified, bit2 certified "west" (llum1),
// bit3 certified "south" (llvm1), bit4 exact llum1, bit5 exact llvm1, bit6 exact IsInsideQuadrangle of the
// exact new position.  tests assert bit0 => bits 2,3 == bits 4,5 and bit0 & bit1 => bit6.
template <int UV>
__global__ void __launch_bounds__(ST_BLOCK)
k_cert_selftest(const AdvectGrid g, long long n, const pt* __restrict__ yx, const int2* __restrict__ cell,
                const float4* __restrict__ vel, uint8_t* __restrict__ flags)
{
    const long long p = (long long)blockIdx.x * ST_BLOCK + threadIdx.x;
    if (p >= n) return;
    const pt P = yx[p];
    const int2 cc = cell[p];
    const int Ni = g.Ni, c = cc.x * Ni + cc.y;
    const float4 w = vel[p];                                       // uL, uR, vB, vT
    const float kdt = __double2float_rn(g.rdt / 1000.0);
    pt Pn;
    const CertLane cl = cert_eval<UV>(g.frames[2 * (size_t)c], g.frames[2 * (size_t)c + 1], P,
                                      w.x, w.y, w.z, w.w, g.rdt, kdt, Pn);
    const pt bl = g.F[c - Ni - 1], br = g.F[c - Ni], ul = g.F[c - 1], ur = g.F[c];
    bool llum1 = false, llvm1 = false;
    double zU, zV;
    if (UV == 1) {
        llum1 = intersect2seg(P, ur, g.V[c - Ni], g.V[c]);
        llvm1 = intersect2seg(P, ur, g.U[c - 1], g.U[c]);
        zU = (double)(llum1 ? w.x : w.y); zV = (double)(llvm1 ? w.z : w.w);
    } else {
        zU = __dmul_rn(0.5, __dadd_rn((double)w.y, (double)w.x));
        zV = __dmul_rn(0.5, __dadd_rn((double)w.w, (double)w.z));
    }
    pt Pe;
    Pe.x = __dadd_rn(P.x, __ddiv_rn(__dmul_rn(zU, g.rdt), 1000.));
    Pe.y = __dadd_rn(P.y, __ddiv_rn(__dmul_rn(zV, g.rdt), 1000.));
    const bool in = inside_quad(Pe.y, Pe.x, bl, br, ur, ul);
    flags[p] = (uint8_t)((int)cl.pick_ok | ((int)cl.in_ok << 1) | ((int)cl.left << 2) | ((int)cl.below << 3) |
                         ((int)llum1 << 4) | ((int)llvm1 << 5) | ((int)in << 6));
}
}  // namespace st
