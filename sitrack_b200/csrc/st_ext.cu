// st_ext.cu -- k_advect_ext: optional physics beyond the reference, off by default.
//
// The reference advances a buoy with ONE Euler step per record from a face velocity picked by two
// segment tests, and follows it across at most ONE cell boundary (si3_part_tracker.py:423-484).  This
// kernel keeps the same state, records, kill rules and outputs but lets the caller choose
//   scheme   1 Euler | 2 midpoint Runge-Kutta | 4 classical Runge-Kutta (the record's field is frozen
//            during the step, so the stages differ in space only);
//   interp   0 the reference's face pick (iUVstrategy) | 1 C-grid linear: u varies linearly between the
//            west and east U-points of the cell, v between the south and north V-points;
//   max_hops how many cell boundaries a step (or a stage) may cross: orientation walk -- move over every
//            edge the target lies outside of (both edges of a corner at once), repeat.
// None of this can be checked against the reference (it has no such modes): tests/test_ext_physics.py
// checks it against closed forms (exact amplification factors of each scheme on a linear field, multi-cell
// walks against a direct search) and against the bit-exact step where the two must coincide
// (scheme 1 / interp 0 / max_hops 1 on steps that stay in or leave through one edge).
#include "st_kernels.h"

namespace st {

struct ExtPhysics { int scheme, interp, max_hops; };

// > 0 when C is to the left of A->B (anticlockwise), plain FP64 with FMA allowed: no parity contract here
__device__ __forceinline__ double orient(pt A, pt B, pt C)
{
    return (B.x - A.x) * (C.y - A.y) - (B.y - A.y) * (C.x - A.x);
}

// Orientation walk: cell (jT,iT) -> the cell that contains P, at most max_hops moves.  Cells are kept
// inside the band where every stencil access of the step is in range (Survive's test 1, 2 <= j <= Nj-3).
// Returns true when P is inside the final cell.
__device__ __forceinline__ bool walk_to(const AdvectGrid& g, pt P, int& jT, int& iT, int max_hops)
{
    const int Ni = g.Ni;
    for (int h = 0; ; ++h) {
        const int c = jT * Ni + iT;
        const pt bl = ldg_pt(g.F, c - Ni - 1), br = ldg_pt(g.F, c - Ni);
        const pt ul = ldg_pt(g.F, c - 1),      ur = ldg_pt(g.F, c);
        const int dj = (int)(orient(ur, ul, P) < 0.0) - (int)(orient(bl, br, P) < 0.0);
        const int di = (int)(orient(br, ur, P) < 0.0) - (int)(orient(ul, bl, P) < 0.0);
        if ((dj | di) == 0) return true;
        if (h >= max_hops) return false;
        const int jn = min(max(jT + dj, 1), g.Nj - 2), in = min(max(iT + di, 1), g.Ni - 2);
        if (jn == jT && in == iT) return false;                 // pinned at the rim of the domain
        jT = jn; iT = in;
    }
}

// ice velocity [m/s] seen by a buoy at P in cell (jT,iT)
template <int UV>
__device__ __forceinline__ void velocity_at(const AdvectGrid& g, const float* __restrict__ u,
                                            const float* __restrict__ v, pt P, int jT, int iT, int interp,
                                            double& zU, double& zV)
{
    const int Ni = g.Ni;
    const int c = jT * Ni + iT;
    const double uW = (double)__ldg(u + c - 1), uE = (double)__ldg(u + c);
    const double vS = (double)__ldg(v + c - Ni), vN = (double)__ldg(v + c);
    if (interp == 1) {
        const pt pw = ldg_pt(g.U, c - 1), pe = ldg_pt(g.U, c);
        const pt ps = ldg_pt(g.V, c - Ni), pn = ldg_pt(g.V, c);
        const double ex = pe.x - pw.x, ey = pe.y - pw.y, nx = pn.x - ps.x, ny = pn.y - ps.y;
        double xi = ((P.x - pw.x) * ex + (P.y - pw.y) * ey) / (ex * ex + ey * ey);
        double et = ((P.x - ps.x) * nx + (P.y - ps.y) * ny) / (nx * nx + ny * ny);
        xi = fmin(fmax(xi, 0.0), 1.0); et = fmin(fmax(et, 0.0), 1.0);
        zU = uW + xi * (uE - uW);
        zV = vS + et * (vN - vS);
    } else if (UV == 1) {
        const pt ur = ldg_pt(g.F, c);
        const bool llum1 = intersect2seg(P, ur, ldg_pt(g.V, c - Ni), ldg_pt(g.V, c));     // si3_part_tracker.py:430
        const bool llvm1 = intersect2seg(P, ur, ldg_pt(g.U, c - 1), ldg_pt(g.U, c));      // :431
        zU = llum1 ? uW : uE;
        zV = llvm1 ? vS : vN;
    } else {
        zU = 0.5 * (uE + uW);
        zV = 0.5 * (vN + vS);
    }
}

template <int UV, bool WIN>
__global__ void __launch_bounds__(ST_BLOCK)
k_advect_ext(const AdvectGrid g, const float* __restrict__ u, const float* __restrict__ v,
             const float* __restrict__ ic, BuoyState s, int jrec, StepOut o, ExtPhysics ph)
{
    const long long p = (long long)blockIdx.x * ST_BLOCK + threadIdx.x;
    const bool valid = p < s.nP;
    int8_t al = 0; pt P = {ST_FILL, ST_FILL}; int2 c = make_int2(2, 2);
    if (valid) { al = s.alive[p]; P = ld_stream_pt(s.pos + p); c = s.cell[p]; }
    if (o.n_alive) {
        const int cnt = __syncthreads_count(al == 1);
        if (threadIdx.x == 0 && cnt) atomicAdd(o.n_alive, (unsigned long long)cnt);
    }
    bool active = valid && al == 1, prestart = false;
    if (WIN && active) {
        const int f = s.rec_first[p], l = s.rec_last[p];
        prestart = (jrec + 1 == f);
        active = (jrec >= f) && (jrec <= l);
    }
    pt outp = {ST_FILL, ST_FILL};
    int8_t m = 0;
    if (active) {
        const double h = g.rdt / 1000.0;                        // km per (m/s) over one record
        int jT = c.x, iT = c.y;
        double k1u, k1v;
        velocity_at<UV>(g, u, v, P, jT, iT, ph.interp, k1u, k1v);
        double du = k1u, dv = k1v;                              // the step's mean velocity
        if (ph.scheme == 2) {
            pt Q = {P.y + 0.5 * h * k1v, P.x + 0.5 * h * k1u};
            int j2 = jT, i2 = iT;
            walk_to(g, Q, j2, i2, ph.max_hops);
            velocity_at<UV>(g, u, v, Q, j2, i2, ph.interp, du, dv);
        } else if (ph.scheme == 4) {
            double k2u, k2v, k3u, k3v, k4u, k4v;
            int js = jT, is = iT;
            pt Q = {P.y + 0.5 * h * k1v, P.x + 0.5 * h * k1u};
            walk_to(g, Q, js, is, ph.max_hops);
            velocity_at<UV>(g, u, v, Q, js, is, ph.interp, k2u, k2v);
            Q.y = P.y + 0.5 * h * k2v; Q.x = P.x + 0.5 * h * k2u;
            walk_to(g, Q, js, is, ph.max_hops);
            velocity_at<UV>(g, u, v, Q, js, is, ph.interp, k3u, k3v);
            Q.y = P.y + h * k3v; Q.x = P.x + h * k3u;
            walk_to(g, Q, js, is, ph.max_hops);
            velocity_at<UV>(g, u, v, Q, js, is, ph.interp, k4u, k4v);
            du = (k1u + 2.0 * k2u + 2.0 * k3u + k4u) * (1.0 / 6.0);
            dv = (k1v + 2.0 * k2v + 2.0 * k3v + k4v) * (1.0 / 6.0);
        }
        pt Pn = {P.y + h * dv, P.x + h * du};
        outp = Pn; m = 1;                                       // recorded before any kill decision (:459-460)
        // follow the buoy into its new host cell; every cell it enters must pass Survive (tracking.py:62-93)
        const int Ni = g.Ni;
        int8_t a2 = 1;
        for (int hop = 0; hop < ph.max_hops; ++hop) {
            const int cc = jT * Ni + iT;
            const pt bl = ldg_pt(g.F, cc - Ni - 1), br = ldg_pt(g.F, cc - Ni);
            const pt ul = ldg_pt(g.F, cc - 1),      ur = ldg_pt(g.F, cc);
            const int dj = (int)(orient(ur, ul, Pn) < 0.0) - (int)(orient(bl, br, Pn) < 0.0);
            const int di = (int)(orient(br, ur, Pn) < 0.0) - (int)(orient(ul, bl, Pn) < 0.0);
            if ((dj | di) == 0) break;
            jT += dj; iT += di;
            if (killed(jT, iT, g.Nj, Ni, g.tmask, ic, g.rmin_conc)) { a2 = 0; break; }
        }
        st_stream_pt(s.pos + p, Pn);
        if (!a2) jT |= ST_DEAD_BIT;
        if (jT != c.x || iT != c.y) s.cell[p] = make_int2(jT, iT);
        if (!a2) s.alive[p] = 0;
    } else if (WIN && prestart) {
        outp = P; m = 1;
    }
    if (valid) {
        if (o.yx) put_row_yx(o, p, outp);
        if (o.mask) __stcs(o.mask + p, m);
        if (o.latlon) put_row_pt(o.latlon, p, inv_stere(outp, g.proj), o.f4);
    }
}

cudaError_t launch_advect_ext(const AdvectGrid& g, const float* u, const float* v, const float* ic,
                              const BuoyState& s, int jrec, const StepOut& o, int scheme, int interp, int max_hops,
                              cudaStream_t st)
{
    if (s.nP <= 0) return cudaSuccess;
    const bool win = s.rec_first != nullptr;
    const dim3 grid((unsigned)((s.nP + ST_BLOCK - 1) / ST_BLOCK)), block(ST_BLOCK);
    const ExtPhysics ph{scheme, interp, max_hops};
    if (g.uv_strategy == 1) {
        if (win) k_advect_ext<1, true><<<grid, block, 0, st>>>(g, u, v, ic, s, jrec, o, ph);
        else     k_advect_ext<1, false><<<grid, block, 0, st>>>(g, u, v, ic, s, jrec, o, ph);
    } else {
        if (win) k_advect_ext<0, true><<<grid, block, 0, st>>>(g, u, v, ic, s, jrec, o, ph);
        else     k_advect_ext<0, false><<<grid, block, 0, st>>>(g, u, v, ic, s, jrec, o, ph);
    }
    return cudaGetLastError();
}

}  // namespace st
