// st_experiments.cuh -- launchers of the round-1 experiment kernels (variants 4, 6-11), compiled only with
// -DST_EXPERIMENTS (python -m sitrack_b200.build -DST_EXPERIMENTS=1 -o lib_exp.so).  They are bit-identical to the
// default kernel and slower (profiles/README.md); kept for A/B measurements, not part of the product build.
#pragma once
namespace st {
static cudaError_t launch_advect_experiment(const AdvectGrid& g, const float* u, const float* v, const float* ic,
                                            const BuoyState& s, int jrec, const StepOut& o, int variant, cudaStream_t st)
{
    const bool win = s.rec_first != nullptr;
    // variants 12 and 9: the one-block-per-tile form of the tuned step (k_advect_step)
#define ST_LAUNCH(BLK_, MINB_)                                                                              \
    do {                                                                                                    \
        const dim3 gr((unsigned)((s.nP + BLK_ - 1) / BLK_)), bl(BLK_);                                       \
        if (g.uv_strategy == 1) {                                                                           \
            if (win) k_advect_step<1, true, BLK_, MINB_><<<gr, bl, 0, st>>>(g, u, v, ic, s, jrec, o);        \
            else     k_advect_step<1, false, BLK_, MINB_><<<gr, bl, 0, st>>>(g, u, v, ic, s, jrec, o);       \
        } else {                                                                                            \
            if (win) k_advect_step<0, true, BLK_, MINB_><<<gr, bl, 0, st>>>(g, u, v, ic, s, jrec, o);        \
            else     k_advect_step<0, false, BLK_, MINB_><<<gr, bl, 0, st>>>(g, u, v, ic, s, jrec, o);       \
        }                                                                                                   \
    } while (0)
    // variants 6, 7, 10, 11: persistent CTAs with a CTA-wide cross-tile walk queue (st_persist.cuh)
    if (variant == 6 || variant == 7 || variant == 10 || variant == 11) {
        int dev = 0;
        cudaGetDevice(&dev);
        static int sm_of[64] = {0};
        if (!sm_of[dev & 63]) cudaDeviceGetAttribute(&sm_of[dev & 63], cudaDevAttrMultiProcessorCount, dev);
        const int n_sm = sm_of[dev & 63];
#define ST_PERSIST(BLK_, MINB_)                                                                             \
        do {                                                                                                \
            const int ntiles = (int)((s.nP + BLK_ - 1) / BLK_);                                              \
            const int nblk = ntiles < MINB_ * n_sm ? ntiles : MINB_ * n_sm;                                  \
            if (g.uv_strategy == 1) {                                                                       \
                if (win) k_advect_persist<1, true, BLK_, MINB_><<<nblk, BLK_, 0, st>>>(g, u, v, ic, s, jrec, o, ntiles);  \
                else     k_advect_persist<1, false, BLK_, MINB_><<<nblk, BLK_, 0, st>>>(g, u, v, ic, s, jrec, o, ntiles); \
            } else {                                                                                        \
                if (win) k_advect_persist<0, true, BLK_, MINB_><<<nblk, BLK_, 0, st>>>(g, u, v, ic, s, jrec, o, ntiles);  \
                else     k_advect_persist<0, false, BLK_, MINB_><<<nblk, BLK_, 0, st>>>(g, u, v, ic, s, jrec, o, ntiles); \
            }                                                                                               \
        } while (0)
        if (variant == 7) ST_PERSIST(128, 8);
        else if (variant == 10) ST_PERSIST(64, 18);
        else if (variant == 11) ST_PERSIST(64, 20);
        else ST_PERSIST(64, 16);
#undef ST_PERSIST
        return cudaGetLastError();
    }
    if (variant == 8) {
        int dev = 0, n_sm = 148;
        cudaGetDevice(&dev);
        static int sm_of[64] = {0};
        if (!sm_of[dev & 63]) cudaDeviceGetAttribute(&sm_of[dev & 63], cudaDevAttrMultiProcessorCount, dev);
        n_sm = sm_of[dev & 63];
        const int ntiles = (int)((s.nP + PIPE_BLK - 1) / PIPE_BLK);
        const int nblk = ntiles < 3 * n_sm ? ntiles : 3 * n_sm;
        const size_t smem = sizeof(PipeSmem);
#define ST_PIPE(UV_, WIN_)                                                                                   \
        do {                                                                                                 \
            static bool attr[64] = {false};                                                                  \
            if (!attr[dev & 63]) { cudaFuncSetAttribute(k_advect_pipe<UV_, WIN_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr[dev & 63] = true; } \
            k_advect_pipe<UV_, WIN_><<<nblk, PIPE_BLK, smem, st>>>(g, u, v, ic, s, jrec, o, ntiles);         \
        } while (0)
        if (g.uv_strategy == 1) { if (win) ST_PIPE(1, true); else ST_PIPE(1, false); }
        else                    { if (win) ST_PIPE(0, true); else ST_PIPE(0, false); }
#undef ST_PIPE
        return cudaGetLastError();
    }
    switch (variant) {
    case 9: ST_LAUNCH(256, 4); break;       // one block per tile, 256 threads
    default: ST_LAUNCH(128, 10); break;     // variant 12: one block per tile, 128 threads x 10 blocks/SM
    }
#undef ST_LAUNCH
    return cudaGetLastError();
}
}  // namespace st
