// st_persist.cuh -- k_advect_persist: the tuned step as persistent CTAs with a cross-tile walk queue.
//
// In k_advect_step the dense walk pass runs once per 128-buoy block with ~20 live threads while the
// block's other warps are already done: the block holds its registers through three dependent
// load rounds (cell corners -> outward neighbours -> tmask/siconc stencil).  An A/B build without the
// walk runs in 201 us instead of 308 us, i.e. a third of the kernel is this tail.  Here a CTA loops
// over tiles, keeps the queue of crossings in shared memory ACROSS tiles and walks BLK entries at a
// time with every thread busy; the next tile's state is loaded into registers while the current one
// is computed.  Arithmetic and results are those of k_advect_step, bit for bit.
#pragma once
#include "st_kernels.h"

namespace st {

template <int UV, bool WIN, int BLK, int MINB>
__global__ void __launch_bounds__(BLK, MINB)
k_advect_persist(const AdvectGrid g, const float* __restrict__ u, const float* __restrict__ v,
                 const float* __restrict__ ic, BuoyState s, int jrec, StepOut o, int ntiles)
{
    constexpr int QCAP = 2 * BLK;
    __shared__ pt qP[QCAP], qPn[QCAP];
    __shared__ int2 qC[QCAP];
    __shared__ unsigned qI[QCAP];
    __shared__ int sCnt[2][BLK / 32];
    __shared__ int sAlive;

    const int tid = threadIdx.x;
    if (tid == 0) sAlive = 0;

    auto load_state = [&](int tile, int8_t& al, pt& P, int2& c2) {
        const long long p = (long long)tile * BLK + tid;
        al = 0; P.y = ST_FILL; P.x = ST_FILL; c2 = make_int2(2, 2);
        if (tile < ntiles && p < s.nP) {
            al = __ldcs(s.alive + p);
            P = ld_stream_pt(s.pos + p);
            c2 = __ldcs(s.cell + p);
        }
    };
    auto walk_pass = [&](int lo, int n) {
#ifdef ST_X_NOWALK
        n = 0;                                   // timing experiment only: results are wrong
#endif
        if (tid < n) {
            const int e = lo + tid;
            int2 cc = qC[e];
            const int j0 = cc.x, i0 = cc.y;
            int8_t a2 = 1;
            pt A, B;
            { const double2 q = *reinterpret_cast<const double2*>(&qP[e]); A.y = q.x; A.x = q.y; }
            { const double2 q = *reinterpret_cast<const double2*>(&qPn[e]); B.y = q.x; B.x = q.y; }
            walk_cell(g, ic, A, B, cc.x, cc.y, a2);
            const unsigned p = qI[e];
            if (!a2) cc.x |= ST_DEAD_BIT;
            if (cc.x != j0 || cc.y != i0) __stcs(s.cell + p, cc);
            if (!a2) s.alive[p] = 0;
        }
    };

    int8_t al; pt P; int2 c2;
    load_state((int)blockIdx.x, al, P, c2);
    int qn = 0, my_alive = 0, it = 0;
    for (int tile = (int)blockIdx.x; tile < ntiles; tile += (int)gridDim.x, ++it) {
        const long long p = (long long)tile * BLK + tid;
        const bool valid = p < s.nP;
        // next tile's state: in flight while this tile is computed
        int8_t nal; pt nP_; int2 nc2;
        load_state(tile + (int)gridDim.x, nal, nP_, nc2);

        my_alive += (al == 1);
        bool active = valid && al == 1;
        bool prestart = false;
        if (WIN && active) {
            const int f = s.rec_first[p], l = s.rec_last[p];
            prestart = (jrec + 1 == f);
            active = (jrec >= f) && (jrec <= l);
        }
        const int2 cw = active ? c2 : make_int2(2, 2);
        const int Ni = g.Ni;
        const int c = cw.x * Ni + cw.y;
        const pt bl = ldg_pt(g.F, c - Ni - 1), br = ldg_pt(g.F, c - Ni);
        const pt ul = ldg_pt(g.F, c - 1),      ur = ldg_pt(g.F, c);
        double zU, zV;
        if (UV == 1) {
            const pt v0 = ldg_pt(g.V, c - Ni), v1 = ldg_pt(g.V, c);
            const pt u0 = ldg_pt(g.U, c - 1),  u1 = ldg_pt(g.U, c);
            const float uL = __ldg(u + c - 1), uR = __ldg(u + c);
            const float vB = __ldg(v + c - Ni), vT = __ldg(v + c);
            const int cb = __ldg(g.cellbits + c);                 // orientation of (ur,v0,v1) and (ur,u0,u1)
            const bool llum1 = intersect2seg_pre(P, ur, v0, v1, cb & 1);      // si3_part_tracker.py:430
            const bool llvm1 = intersect2seg_pre(P, ur, u0, u1, cb & 2);      // :431
            zU = (double)(llum1 ? uL : uR);
            zV = (double)(llvm1 ? vB : vT);
        } else {
            zU = __dmul_rn(0.5, __dadd_rn((double)__ldg(u + c), (double)__ldg(u + c - 1)));
            zV = __dmul_rn(0.5, __dadd_rn((double)__ldg(v + c), (double)__ldg(v + c - Ni)));
        }
        pt Pn;
        Pn.x = __dadd_rn(P.x, div1000(__dmul_rn(zU, g.rdt)));      // :452-458
        Pn.y = __dadd_rn(P.y, div1000(__dmul_rn(zV, g.rdt)));
        const bool in = inside_quad2(Pn.y, Pn.x, bl, br, ur, ul, active);
        const bool cross = active && !in;
        pt outp = {ST_FILL, ST_FILL};
        int8_t m = 0;
        if (active) {
            outp = Pn; m = 1;
            st_stream_pt(s.pos + p, outp);
        } else if (WIN && prestart) {
            outp = P; m = 1;
        }
        if (valid) {
            if (o.yx) put_row_yx(o, p, outp);
            if (o.mask) __stcs(o.mask + p, m);
            if (o.latlon) {
                pt ll; ll.y = g.proj.fill_lat; ll.x = g.proj.fill_lon;            // :493 on a fill row
                if (m) ll = inv_stere_fast(outp, g.proj, g.atab);
                put_row_pt(o.latlon, p, ll, o.f4);
            }
        }
        // append this tile's crossings to the queue (deterministic order: warp, then lane)
        const unsigned bal = __ballot_sync(0xffffffffu, cross);
        const int par = it & 1;
        if ((tid & 31) == 0) sCnt[par][tid >> 5] = __popc(bal);
        __syncthreads();                                          // B1: counts visible, earlier pass finished
        int base = qn, tot = 0;
#pragma unroll
        for (int w = 0; w < BLK / 32; ++w) {
            const int cwt = sCnt[par][w];
            if (w < (tid >> 5)) base += cwt;
            tot += cwt;
        }
        if (cross) {
            const int e = base + __popc(bal & ((1u << (tid & 31)) - 1u));
            *reinterpret_cast<double2*>(&qP[e]) = make_double2(P.y, P.x);
            *reinterpret_cast<double2*>(&qPn[e]) = make_double2(outp.y, outp.x);
            qC[e] = c2; qI[e] = (unsigned)p;
        }
        qn += tot;
        if (qn >= BLK) {
            __syncthreads();                                      // B2: entries visible
            qn -= BLK;
            walk_pass(qn, BLK);
        }
        al = nal; P = nP_; c2 = nc2;
    }
    __syncthreads();
    walk_pass(0, qn);                                             // flush (qn < BLK)
    if (o.n_alive) {
        const int wsum = __reduce_add_sync(0xffffffffu, my_alive);
        if ((tid & 31) == 0 && wsum) atomicAdd(&sAlive, wsum);
        __syncthreads();
        if (tid == 0 && sAlive) atomicAdd(o.n_alive, (unsigned long long)sAlive);
    }
}

}  // namespace st
