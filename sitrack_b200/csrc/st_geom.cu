// st_geom.cu -- batched forms of the reference's scalar geometry predicates, on
// explicit coordinates.  They back the function surface the tracker's users
// import (sit.intersect2Seg, sit.IsInsideQuadrangle, sit.CrossedEdge,
// sit.NewHostCell, sit.Survive, sit.Haversine): same device primitives as the
// fused step kernel, one element per thread.
#include "st_kernels.h"

namespace st {

__global__ void k_geom_intersect(long long n, const pt* __restrict__ A, const pt* __restrict__ B,
                                 const pt* __restrict__ C, const pt* __restrict__ D, int8_t* __restrict__ out)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = intersect2seg(A[k], B[k], C[k], D[k]);
}

__global__ void k_geom_inside(long long n, const pt* __restrict__ yx, const pt* __restrict__ quads,
                              int8_t* __restrict__ out)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) {
        const pt* q = quads + 4 * k;
        out[k] = inside_quad(yx[k].y, yx[k].x, q[0], q[1], q[2], q[3]);
    }
}

// ring (n,12): [0..3] = cell vertices BL,BR,UR,UL; then the outward neighbours
// [4]=F[jbl-1,ibl] [5]=F[jbr-1,ibr] [6]=F[jbr,ibr+1] [7]=F[jur,iur+1]
// [8]=F[jul+1,iul] [9]=F[jur+1,iur] [10]=F[jul,iul-1] [11]=F[jbl,ibl-1]
// kcross_in == nullptr: CrossedEdge then NewHostCell; else NewHostCell for the given edge.
__global__ void k_geom_walk(long long n, const pt* __restrict__ p1, const pt* __restrict__ p2,
                            const pt* __restrict__ ring, const int32_t* __restrict__ kcross_in,
                            int32_t* __restrict__ kcross, int32_t* __restrict__ knhc)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const pt* r = ring + 12 * k;
    const pt A = p1[k], B = p2[k];
    const int kc = kcross_in ? kcross_in[k] : crossed_edge(A, B, r[0], r[1], r[2], r[3]);
    int kn = kc;
    if (kc == 1)      { if (intersect2seg(A, B, r[0], r[4]))  kn = 5; else if (intersect2seg(A, B, r[1], r[5]))  kn = 6; }
    else if (kc == 2) { if (intersect2seg(A, B, r[1], r[6]))  kn = 6; else if (intersect2seg(A, B, r[2], r[7]))  kn = 7; }
    else if (kc == 3) { if (intersect2seg(A, B, r[3], r[8]))  kn = 8; else if (intersect2seg(A, B, r[2], r[9]))  kn = 7; }
    else if (kc == 4) { if (intersect2seg(A, B, r[3], r[10])) kn = 8; else if (intersect2seg(A, B, r[0], r[11])) kn = 5; }
    if (kcross) kcross[k] = kc;
    if (knhc) knhc[k] = kn;
}

// Survive on pre-gathered 5-point stencils (tracking.py:62-93); tm5/ic5 order:
// [jT,iT] [jT,iT+1] [jT+1,iT] [jT,iT-1] [jT-1,iT-1].  ic5 == nullptr skips test 3.
__global__ void k_geom_survive(long long n, const int32_t* __restrict__ ji, int Nj, int Ni,
                               const int8_t* __restrict__ tm5, const double* __restrict__ ic5,
                               double rmin_conc, int32_t* __restrict__ out)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int jT = ji[2 * k], iT = ji[2 * k + 1];
    int kill = 0;
    if (jT == 0 || jT == 1 || jT == Nj - 2 || jT == Nj - 1 || iT == 0 || iT == 1 || iT == Ni - 2 || iT == Ni - 1)
        kill = 1;
    if (!kill) {
        const int8_t* t = tm5 + 5 * k;
        if (t[0] + t[1] + t[2] + t[3] + t[4] < 5) kill = 1;
    }
    if (!kill && ic5) {
        const double* c = ic5 + 5 * k;
        double s = __dadd_rn(c[0], c[1]); s = __dadd_rn(s, c[2]); s = __dadd_rn(s, c[3]); s = __dadd_rn(s, c[4]);
        if (__dmul_rn(0.2, s) < rmin_conc) kill = 1;
    }
    out[k] = kill;
}

__global__ void k_haversine(long long n, double plat, double plon, const double* __restrict__ lat,
                            const double* __restrict__ lon, double* __restrict__ out)
{
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) out[k] = haversine_km(plat, plon, lat[k], lon[k]);
}

#define GRID(n) (unsigned)(((n) + 255) / 256), 256

cudaError_t launch_geom_intersect(long long n, const pt* A, const pt* B, const pt* C, const pt* D,
                                  int8_t* out, cudaStream_t st)
{
    if (n > 0) k_geom_intersect<<<GRID(n), 0, st>>>(n, A, B, C, D, out);
    return cudaGetLastError();
}
cudaError_t launch_geom_inside(long long n, const pt* yx, const pt* quads, int8_t* out, cudaStream_t st)
{
    if (n > 0) k_geom_inside<<<GRID(n), 0, st>>>(n, yx, quads, out);
    return cudaGetLastError();
}
cudaError_t launch_geom_walk(long long n, const pt* p1, const pt* p2, const pt* ring, const int32_t* kcross_in,
                             int32_t* kcross, int32_t* knhc, cudaStream_t st)
{
    if (n > 0) k_geom_walk<<<GRID(n), 0, st>>>(n, p1, p2, ring, kcross_in, kcross, knhc);
    return cudaGetLastError();
}
cudaError_t launch_geom_survive(long long n, const int32_t* ji, int Nj, int Ni, const int8_t* tm5,
                                const double* ic5, double rmin_conc, int32_t* out, cudaStream_t st)
{
    if (n > 0) k_geom_survive<<<GRID(n), 0, st>>>(n, ji, Nj, Ni, tm5, ic5, rmin_conc, out);
    return cudaGetLastError();
}
cudaError_t launch_haversine(long long n, double plat, double plon, const double* lat, const double* lon,
                             double* out, cudaStream_t st)
{
    if (n > 0) k_haversine<<<GRID(n), 0, st>>>(n, plat, plon, lat, lon, out);
    return cudaGetLastError();
}

}  // namespace st
