// st_pipe.cuh -- k_advect_pipe: the persistent, software-pipelined form of the per-record step.
//
// The tuned one-shot kernel (k_advect_step) is latency-bound on B200: the profile shows 39 % of
// warp stalls on the long scoreboard (13 % waiting for the buoy-state lines from HBM, 15 % for the
// geometry gathers, 5 % in the cell-walk tail) at 51 % issue utilisation.  This kernel removes the
// exposed latency instead of adding occupancy:
//   * persistent CTAs (3 per SM) loop over 256-buoy tiles;
//   * the state tile (pos / cell / alive: 4096 + 2048 + 256 B) arrives through a 3-stage ring of
//     TMA bulk copies (cp.async.bulk ... mbarrier::complete_tx), issued two tiles ahead by one thread;
//   * each thread gathers its NEXT buoy's 8 geometry points and 4 velocities with cp.async
//     (LDGSTS, no registers) into its own shared-memory slots while it computes the current one;
//   * cell walks are queued in shared memory ACROSS tiles and processed 128 at a time by four
//     full warps, so the divergent slow path never runs with 3-4 live lanes nor holds a block tail.
// Results are bit-identical to k_advect_step / k_advect_step_v1 (tests/test_gpu_parity.py).
#pragma once
#include "st_kernels.h"

namespace st {

constexpr int PIPE_BLK = 256;        // buoys per tile = threads per CTA
constexpr int PIPE_NS = 3;           // state ring depth
constexpr int PIPE_QT = 128;         // dense walk pass size
constexpr int PIPE_QCAP = PIPE_QT + PIPE_BLK;
constexpr uint32_t PIPE_TILE_BYTES = PIPE_BLK * (sizeof(pt) + sizeof(int2) + 1);

struct PipeSmem {
    pt pos[PIPE_NS][PIPE_BLK];
    int2 cell[PIPE_NS][PIPE_BLK];
    int8_t alive[PIPE_NS][PIPE_BLK];
    pt geo[8][PIPE_BLK];             // bl br ul ur v0 v1 u0 u1 of the thread's next buoy
    float4 vel[PIPE_BLK];            // uL uR vB vT
    pt qP[PIPE_QCAP], qPn[PIPE_QCAP];
    int2 qC[PIPE_QCAP];
    unsigned qI[PIPE_QCAP];          // buoy index within this rank
    unsigned long long full[PIPE_NS];
    int wcnt[2][PIPE_BLK / 32];
    int nalive;
};

__device__ __forceinline__ void cp_async16(void* dst, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(void* dst, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
}

// issue the gathers of one buoy (cell c2) into the calling thread's slots
template <int UV>
__device__ __forceinline__ void pipe_gather(PipeSmem& sm, int tid, const AdvectGrid& g, const float* __restrict__ u,
                                            const float* __restrict__ v, int2 c2)
{
    const int Ni = g.Ni;
    const int c = c2.x * Ni + c2.y;
    cp_async16(&sm.geo[0][tid], g.F + (c - Ni - 1));
    cp_async16(&sm.geo[1][tid], g.F + (c - Ni));
    cp_async16(&sm.geo[2][tid], g.F + (c - 1));
    cp_async16(&sm.geo[3][tid], g.F + c);
    if (UV == 1) {
        cp_async16(&sm.geo[4][tid], g.V + (c - Ni));
        cp_async16(&sm.geo[5][tid], g.V + c);
        cp_async16(&sm.geo[6][tid], g.U + (c - 1));
        cp_async16(&sm.geo[7][tid], g.U + c);
    }
    float* vs = reinterpret_cast<float*>(&sm.vel[tid]);
    cp_async4(vs + 0, u + (c - 1));
    cp_async4(vs + 1, u + c);
    cp_async4(vs + 2, v + (c - Ni));
    cp_async4(vs + 3, v + c);
}

__device__ __forceinline__ pt lds_pt(const pt* a)
{
    const double2 q = *reinterpret_cast<const double2*>(a);
    pt r; r.y = q.x; r.x = q.y; return r;
}

// one dense pass over queue entries [lo, lo+n)
__device__ __forceinline__ void pipe_walk(PipeSmem& sm, int tid, int lo, int n, const AdvectGrid& g,
                                          const float* __restrict__ ic, const BuoyState& s)
{
    if (tid < n) {
        const int e = lo + tid;
        int2 cc = sm.qC[e];
        const int j0 = cc.x, i0 = cc.y;
        int8_t a2 = 1;
        walk_cell(g, ic, lds_pt(&sm.qP[e]), lds_pt(&sm.qPn[e]), cc.x, cc.y, a2);
        const unsigned p = sm.qI[e];
        if (!a2) cc.x |= ST_DEAD_BIT;
        if (cc.x != j0 || cc.y != i0) __stcs(s.cell + p, cc);
        if (!a2) s.alive[p] = 0;
    }
}

template <int UV, bool WIN>
__global__ void __launch_bounds__(PIPE_BLK, 3)
k_advect_pipe(const AdvectGrid g, const float* __restrict__ u, const float* __restrict__ v,
              const float* __restrict__ ic, BuoyState s, int jrec, StepOut o, int ntiles)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    PipeSmem& sm = *reinterpret_cast<PipeSmem*>(smem_raw);
    const int tid = threadIdx.x;
    const int nloc = (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;   // tiles of this CTA

    if (tid == 0) {
        for (int i = 0; i < PIPE_NS; ++i) mbar_init(&sm.full[i], 1);
        sm.nalive = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue_tile = [&](int k) {                       // thread 0 only: TMA of local tile k into stage k%NS
        const int st = k % PIPE_NS;
        const long long p0 = ((long long)blockIdx.x + (long long)k * gridDim.x) * PIPE_BLK;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(&sm.full[st], PIPE_TILE_BYTES);
        tma_load_1d(sm.pos[st], s.pos + p0, PIPE_BLK * sizeof(pt), &sm.full[st]);
        tma_load_1d(sm.cell[st], s.cell + p0, PIPE_BLK * sizeof(int2), &sm.full[st]);
        tma_load_1d(sm.alive[st], s.alive + p0, PIPE_BLK, &sm.full[st]);
    };
    if (tid == 0)
        for (int k = 0; k < PIPE_NS && k < nloc; ++k) issue_tile(k);

    // state of the current tile in registers; gathers of tile 0
    int8_t al = 0; pt P = {ST_FILL, ST_FILL}; int2 c2 = make_int2(2, 2);
    bool active = false, prestart = false;
    auto read_state = [&](int k) {
        const int st = k % PIPE_NS;
        mbar_wait(&sm.full[st], (uint32_t)((k / PIPE_NS) & 1));
        const long long p = ((long long)blockIdx.x + (long long)k * gridDim.x) * PIPE_BLK + tid;
        al = sm.alive[st][tid];
        P = lds_pt(&sm.pos[st][tid]);
        c2 = sm.cell[st][tid];
        active = (p < s.nP) && al == 1;
        prestart = false;
        if (WIN && active) {
            const int f = s.rec_first[p], l = s.rec_last[p];
            prestart = (jrec + 1 == f);
            active = (jrec >= f) && (jrec <= l);
        }
    };
    if (nloc > 0) {
        read_state(0);
        if (active) pipe_gather<UV>(sm, tid, g, u, v, c2);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }

    int qn = 0;                                          // queue length, uniform across the CTA
    int my_alive = 0;
    for (int k = 0; k < nloc; ++k) {
        const long long p0 = ((long long)blockIdx.x + (long long)k * gridDim.x) * PIPE_BLK;
        const long long p = p0 + tid;
        const bool valid = p < s.nP;
        my_alive += (valid && al == 1);
        // ---- operands of this buoy out of shared memory ------------------------------------
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        pt bl, br, ul, ur, v0, v1, u0, u1; float4 vel = make_float4(0.f, 0.f, 0.f, 0.f);
        bl = br = ul = ur = v0 = v1 = u0 = u1 = P;
        if (active) {
            bl = lds_pt(&sm.geo[0][tid]); br = lds_pt(&sm.geo[1][tid]);
            ul = lds_pt(&sm.geo[2][tid]); ur = lds_pt(&sm.geo[3][tid]);
            if (UV == 1) {
                v0 = lds_pt(&sm.geo[4][tid]); v1 = lds_pt(&sm.geo[5][tid]);
                u0 = lds_pt(&sm.geo[6][tid]); u1 = lds_pt(&sm.geo[7][tid]);
            }
            vel = sm.vel[tid];
        }
        const pt Pcur = P; const int2 ccur = c2; const bool act = active, pre = prestart;
        pt outp = {ST_FILL, ST_FILL};
        int8_t m = 0;
        bool cross = false;
        if (act) {
            double zU, zV;
            if (UV == 1) {
                const bool llum1 = intersect2seg(Pcur, ur, v0, v1);      // si3_part_tracker.py:430
                const bool llvm1 = intersect2seg(Pcur, ur, u0, u1);      // :431
                zU = (double)(llum1 ? vel.x : vel.y);
                zV = (double)(llvm1 ? vel.z : vel.w);
            } else {
                zU = __dmul_rn(0.5, __dadd_rn((double)vel.y, (double)vel.x));
                zV = __dmul_rn(0.5, __dadd_rn((double)vel.w, (double)vel.z));
            }
            outp.x = __dadd_rn(Pcur.x, div1000(__dmul_rn(zU, g.rdt)));    // :452-458
            outp.y = __dadd_rn(Pcur.y, div1000(__dmul_rn(zV, g.rdt)));
            m = 1;
        } else if (WIN && pre) {
            outp = Pcur; m = 1;
        }
        // ---- this thread's slots are consumed: prefetch the next tile's operands --------------
        if (k + 1 < nloc) {
            read_state(k + 1);
            if (active) pipe_gather<UV>(sm, tid, g, u, v, c2);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        // ---- inside test, rows, state ----------------------------------------------------------
        if (act) {
            cross = !inside_quad_div(outp.y, outp.x, bl, br, ur, ul);
            st_stream_pt(s.pos + p, outp);
        }
        if (valid) {
            if (o.yx) put_row_yx(o, p, outp);
            if (o.mask) __stcs(o.mask + p, m);
            if (o.latlon) {
                pt ll; ll.y = g.proj.fill_lat; ll.x = g.proj.fill_lon;
                if (m) ll = inv_stere_fast(outp, g.proj, g.atab);
                put_row_pt(o.latlon, p, ll, o.f4);
            }
        }
        // ---- queue the crossings of the tile ----------------------------------------------------
        const unsigned bal = __ballot_sync(0xffffffffu, cross);
        const int par = k & 1;
        if ((tid & 31) == 0) sm.wcnt[par][tid >> 5] = __popc(bal);
        __syncthreads();                                  // B1: counts visible; stage k%NS and geo slots consumed
        if (tid == 0 && k + PIPE_NS < nloc) issue_tile(k + PIPE_NS);
        int base = qn, tot = 0;
#pragma unroll
        for (int w = 0; w < PIPE_BLK / 32; ++w) {
            const int cw = sm.wcnt[par][w];
            if (w < (tid >> 5)) base += cw;
            tot += cw;
        }
        if (cross) {
            const int e = base + __popc(bal & ((1u << (tid & 31)) - 1u));
            sm.qP[e].y = Pcur.y; sm.qP[e].x = Pcur.x;
            sm.qPn[e].y = outp.y; sm.qPn[e].x = outp.x;
            sm.qC[e] = ccur; sm.qI[e] = (unsigned)p;
        }
        qn += tot;
        if (qn >= PIPE_QT) {
            __syncthreads();                              // B2: queue entries visible
            while (qn >= PIPE_QT) {
                qn -= PIPE_QT;
                pipe_walk(sm, tid, qn, PIPE_QT, g, ic, s);
            }
        }
    }
    __syncthreads();
    pipe_walk(sm, tid, 0, qn, g, ic, s);                  // flush (qn < PIPE_QT)
    if (o.n_alive) {
        if (my_alive) atomicAdd(&sm.nalive, my_alive);
        __syncthreads();
        if (tid == 0 && sm.nalive) atomicAdd(o.n_alive, (unsigned long long)sm.nalive);
    }
}

}  // namespace st
