// st_locate.cu -- seeding: nearest T-point + survive + containing cell (sm_100a).
//
// Replaces the reference's per-buoy whole-grid Haversine scan
// (sitrack/locate.py:222-276, called from tracking.py:120-160) with a
// coarse-bin spatial hash of the T-points and an expanding-ring search that is
// still an EXACT argmin of the same Haversine expression:
//
//   * every T-point is mapped to the unit-sphere polar stereographic plane
//     Q = (cos(lat) / (1 + sin(lat))) * (sin(lon), -cos(lon));
//   * for two sphere points the haversine term obeys the identity
//         h = sin^2(angle/2) = |P-Q|^2 / ((1+|P|^2) (1+|Q|^2)),
//     so once all bins within Chebyshev ring r of the buoy's bin are scanned,
//     every unscanned point has |P-Q| > r*bin, i.e.
//         d > 2R asin( r*bin / sqrt((1+|P|^2)(1+Qmax^2)) )  =: d_bound(r);
//   * the search stops when best_d <= d_bound(r) (argmin proven) or when
//     d_bound(r) exceeds the largest distance NearestPoint could still accept
//     (0.5 * max(resKM) * 1.2^7) -- then the answer is (-1,-1) whatever the argmin.
//   Ties on equal distance resolve to the lower flat index like numpy's argmin.
#include "st_kernels.h"
#include <math.h>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>

namespace st {

__device__ __forceinline__ void plane_of(double lat, double lon, double& px, double& py)
{
    const double D2R = 0.017453292519943295;
    double sp, cp, sl, cl;
    sincos(lat * D2R, &sp, &cp);
    sincos(lon * D2R, &sl, &cl);
    const double r = cp / (1.0 + sp);
    px = r * sl; py = -r * cl;
}

// ---- build ---------------------------------------------------------------------------
// plane coordinates of every T-point and, per block, the partial extrema {xmin, xmax, ymin, ymax, q2max, resmax}
// (part[6 * blockIdx.x ...]); k_bounds_finish folds the partials.  Grid-stride, one block per SM or more.
__device__ __forceinline__ void reduce6(double v[6])
{
    for (int o = 16; o; o >>= 1) {
        v[0] = fmin(v[0], __shfl_xor_sync(~0u, v[0], o)); v[1] = fmax(v[1], __shfl_xor_sync(~0u, v[1], o));
        v[2] = fmin(v[2], __shfl_xor_sync(~0u, v[2], o)); v[3] = fmax(v[3], __shfl_xor_sync(~0u, v[3], o));
        v[4] = fmax(v[4], __shfl_xor_sync(~0u, v[4], o)); v[5] = fmax(v[5], __shfl_xor_sync(~0u, v[5], o));
    }
}
__device__ __forceinline__ void block_reduce6(double v[6], double* __restrict__ out)
{
    __shared__ double sh[6][32];
    reduce6(v);
    const int w = threadIdx.x >> 5, ln = threadIdx.x & 31;
    if (ln == 0) for (int q = 0; q < 6; ++q) sh[q][w] = v[q];
    __syncthreads();
    if (w == 0) {
        const int nw = blockDim.x >> 5;
        for (int q = 0; q < 6; ++q) v[q] = (ln < nw) ? sh[q][ln] : ((q == 0 || q == 2) ? 1e300 : (q < 4 ? -1e300 : 0.0));
        reduce6(v);
        if (ln == 0) for (int q = 0; q < 6; ++q) out[q] = v[q];
    }
}
__global__ void __launch_bounds__(256)
k_plane_bounds(const double* __restrict__ lat, const double* __restrict__ lon,
               const double* __restrict__ res, int n, double* __restrict__ px,
               double* __restrict__ py, double* __restrict__ part /*[6 * gridDim.x]*/)
{
    double v[6] = {1e300, -1e300, 1e300, -1e300, 0.0, 0.0};
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        double x, y; plane_of(lat[k], lon[k], x, y);
        px[k] = x; py[k] = y;
        v[0] = fmin(v[0], x); v[1] = fmax(v[1], x); v[2] = fmin(v[2], y); v[3] = fmax(v[3], y);
        v[4] = fmax(v[4], x * x + y * y);
        if (res) v[5] = fmax(v[5], res[k]);
    }
    block_reduce6(v, part + 6 * blockIdx.x);
}
__global__ void __launch_bounds__(256)
k_bounds_finish(const double* __restrict__ part, int nblk, double* __restrict__ red /*[6]*/)
{
    double v[6] = {1e300, -1e300, 1e300, -1e300, 0.0, 0.0};
    for (int b = threadIdx.x; b < nblk; b += blockDim.x) {
        const double* q = part + 6 * b;
        v[0] = fmin(v[0], q[0]); v[1] = fmax(v[1], q[1]); v[2] = fmin(v[2], q[2]); v[3] = fmax(v[3], q[3]);
        v[4] = fmax(v[4], q[4]); v[5] = fmax(v[5], q[5]);
    }
    block_reduce6(v, red);
}

__device__ __forceinline__ int bin_of(double x, double y, double x0, double y0, double inv_bin, int nbx, int nby)
{
    int bx = (int)floor((x - x0) * inv_bin), by = (int)floor((y - y0) * inv_bin);
    bx = min(max(bx, 0), nbx - 1); by = min(max(by, 0), nby - 1);
    return by * nbx + bx;
}

__global__ void k_bin_count(const double* __restrict__ px, const double* __restrict__ py, int n,
                            double x0, double y0, double inv_bin, int nbx, int nby, int* __restrict__ cnt)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) atomicAdd(cnt + bin_of(px[k], py[k], x0, y0, inv_bin, nbx, nby), 1);
}

__global__ void k_bin_fill(const double* __restrict__ px, const double* __restrict__ py, int n,
                           double x0, double y0, double inv_bin, int nbx, int nby,
                           const int* __restrict__ start, int* __restrict__ cursor, int* __restrict__ pts)
{
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) {
        const int b = bin_of(px[k], py[k], x0, y0, inv_bin, nbx, nby);
        pts[start[b] + atomicAdd(cursor + b, 1)] = k;
    }
}

cudaError_t locate_build(int Nj, int Ni, const double* d_lat, const double* d_lon, const double* d_res,
                         LocateGrid* out, int** owned_start, int** owned_pts, cudaStream_t st)
{
    const int n = Nj * Ni;
    double *px = nullptr, *py = nullptr, *red = nullptr;
    int *cnt = nullptr, *start = nullptr, *pts = nullptr;
    void* tmp = nullptr;
    size_t tmp_bytes = 0;
    const int nblk = 592;                                     // 4 blocks per SM of a B200
    cudaError_t e;
#define CK(x) do { e = (x); if (e != cudaSuccess) goto fail; } while (0)
    CK(cudaMalloc(&px, sizeof(double) * n)); CK(cudaMalloc(&py, sizeof(double) * n));
    CK(cudaMalloc(&red, sizeof(double) * 6 * (nblk + 1)));
    k_plane_bounds<<<nblk, 256, 0, st>>>(d_lat, d_lon, d_res, n, px, py, red + 6);
    k_bounds_finish<<<1, 256, 0, st>>>(red + 6, nblk, red);
    double h[6];
    CK(cudaMemcpyAsync(h, red, sizeof(h), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    {
        const double w = fmax(h[1] - h[0], 1e-12), ht = fmax(h[3] - h[2], 1e-12);
        double bin = sqrt(w * ht * 2.0 / n);                  // ~2 T-points per bin
        int nbx = (int)(w / bin) + 1, nby = (int)(ht / bin) + 1;
        const int nb = nbx * nby;
        CK(cudaMalloc(&cnt, sizeof(int) * (nb + 1))); CK(cudaMalloc(&start, sizeof(int) * (nb + 1)));
        CK(cudaMalloc(&pts, sizeof(int) * n));
        CK(cudaMemsetAsync(cnt, 0, sizeof(int) * (nb + 1), st));
        const int B = 256, G = (n + B - 1) / B;
        k_bin_count<<<G, B, 0, st>>>(px, py, n, h[0], h[2], 1.0 / bin, nbx, nby, cnt);
        // start[0..nb] = exclusive prefix sum of cnt[0..nb] (cnt[nb] = 0): CUB's device-wide scan (setup, once per grid)
        CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, cnt, start, nb + 1, st));
        CK(cudaMalloc(&tmp, tmp_bytes));
        CK(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cnt, start, nb + 1, st));
        CK(cudaMemsetAsync(cnt, 0, sizeof(int) * nb, st));
        k_bin_fill<<<G, B, 0, st>>>(px, py, n, h[0], h[2], 1.0 / bin, nbx, nby, start, cnt, pts);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(st));
        out->Nj = Nj; out->Ni = Ni; out->latT = d_lat; out->lonT = d_lon; out->resKM = d_res;
        out->nbx = nbx; out->nby = nby; out->x0 = h[0]; out->y0 = h[2]; out->bin = bin; out->inv_bin = 1.0 / bin;
        out->q2max = h[4]; out->res_max = h[5]; out->bin_start = start; out->bin_pts = pts;
        *owned_start = start; *owned_pts = pts;
    }
    cudaFree(px); cudaFree(py); cudaFree(red); cudaFree(cnt); cudaFree(tmp);
    return cudaSuccess;
fail:
    cudaFree(px); cudaFree(py); cudaFree(red); cudaFree(cnt); cudaFree(start); cudaFree(pts); cudaFree(tmp);
    return e;
#undef CK
}

// ---- search ----------------------------------------------------------------------------
struct NearestOpt {
    double rd_found_km;     // used when the grid has no resKM (locate.py:227 default path)
    int accept_mode;        // >=0: number of x1.2 growths of the last accepted radius; -1 never; -2 always
};

// best and runner-up (the runner-up feeds the near-tie flag of k_seed_locate)
__device__ __forceinline__ void scan_bin(const LocateGrid& lg, int b, double plat, double plon,
                                         double& best_d, int& best_k, double& sec_d, int& sec_k)
{
    const int s = __ldg(lg.bin_start + b), e = __ldg(lg.bin_start + b + 1);
    for (int q = s; q < e; ++q) {
        const int k = __ldg(lg.bin_pts + q);
        const double d = haversine_km(plat, plon, __ldg(lg.latT + k), __ldg(lg.lonT + k));
        if (d < best_d || (d == best_d && k < best_k)) { sec_d = best_d; sec_k = best_k; best_d = d; best_k = k; }
        else if (d < sec_d || (d == sec_d && k < sec_k)) { sec_d = d; sec_k = k; }
    }
}

// exact argmin of the Haversine distance over the whole grid, or best_k = -1 when
// the nearest point is provably farther than r_accept_max.
// CUDA's sin/cos/asin and numpy's differ in the last ulp, so an argmin or an acceptance decided by less than
// ST_TIE_REL (relative) is not trusted: the search keeps going until everything unscanned is farther than
// best * (1 + ST_TIE_REL), the runner-up is reported, and the host re-evaluates such seeds with numpy (locate.py).
#define ST_TIE_REL 1e-11
__device__ void nearest_hash(const LocateGrid& lg, double plat, double plon, double r_accept_max,
                             double& best_d, int& best_k, double& sec_d, int& sec_k)
{
    double px, py; plane_of(plat, plon, px, py);
    const double scale = sqrt((1.0 + px * px + py * py) * (1.0 + lg.q2max));
    int bx = (int)floor((px - lg.x0) * lg.inv_bin), by = (int)floor((py - lg.y0) * lg.inv_bin);
    bx = min(max(bx, 0), lg.nbx - 1); by = min(max(by, 0), lg.nby - 1);
    best_d = INFINITY; best_k = -1; sec_d = INFINITY; sec_k = -1;
    const int rmax = max(max(bx, lg.nbx - 1 - bx), max(by, lg.nby - 1 - by));
    for (int r = 0; r <= rmax; ++r) {
        const int y0 = by - r, y1 = by + r, x0 = bx - r, x1 = bx + r;
        for (int yy = max(y0, 0); yy <= min(y1, lg.nby - 1); ++yy) {
            if (yy == y0 || yy == y1) {
                for (int xx = max(x0, 0); xx <= min(x1, lg.nbx - 1); ++xx)
                    scan_bin(lg, yy * lg.nbx + xx, plat, plon, best_d, best_k, sec_d, sec_k);
            } else {
                if (x0 >= 0) scan_bin(lg, yy * lg.nbx + x0, plat, plon, best_d, best_k, sec_d, sec_k);
                if (x1 < lg.nbx && x1 != x0) scan_bin(lg, yy * lg.nbx + x1, plat, plon, best_d, best_k, sec_d, sec_k);
            }
        }
        // lower bound on the distance of anything not scanned yet
        const double sh = fmin(r * lg.bin / scale * (1.0 - 1e-9), 1.0);
        const double d_bound = 2. * 6360. * asin(sh);
        if (best_d * (1.0 + ST_TIE_REL) <= d_bound) return;     // argmin proven, and every near-tie has been seen
        if (d_bound > r_accept_max) { if (best_d > r_accept_max) best_k = -1; if (best_k < 0) return; }
    }
}

__device__ __forceinline__ bool accept_nearest(const LocateGrid& lg, const NearestOpt& no, int k, double d)
{
    if (no.accept_mode == -2) return true;
    if (no.accept_mode < 0) return false;
    double rf = lg.resKM ? __dmul_rn(0.5, __ldg(lg.resKM + k)) : no.rd_found_km;   // locate.py:262
    for (int q = 0; q < no.accept_mode; ++q) rf = __dmul_rn(1.2, rf);               // locate.py:268
    return d < rf;                                                                  // locate.py:266
}

// k_seed_locate: the body of SeedInit's loop (tracking.py:120-160), one thread per buoy.
__global__ void __launch_bounds__(128)
k_seed_locate(const LocateGrid lg, const AdvectGrid g, const float* __restrict__ ic0, long long nP,
              const pt* __restrict__ SG, const pt* __restrict__ SC, SeedOut o, NearestOpt no,
              double r_accept_max, int do_survive, int do_cell)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nP) return;
    const pt ll = SG[p];                                      // [lat, lon]
    double d, d2; int k, k2;
    nearest_hash(lg, ll.y, ll.x, r_accept_max, d, k, d2, k2);
    if (o.flag) {                                             // decisions within ST_TIE_REL: the host re-evaluates with numpy
        int8_t fl = 0;
        if (k >= 0 && k2 >= 0 && d2 - d <= ST_TIE_REL * d2) fl |= 1;                 // argmin near-tie
        if (k >= 0 && no.accept_mode >= 0) {
            double rf = lg.resKM ? __dmul_rn(0.5, __ldg(lg.resKM + k)) : no.rd_found_km;
            for (int q = 0; q < no.accept_mode; ++q) rf = __dmul_rn(1.2, rf);
            if (fabs(d - rf) <= ST_TIE_REL * rf) fl |= 2;                            // acceptance near the threshold
        }
        o.flag[p] = fl;
        if (o.second) o.second[p] = (fl & 1) ? k2 : -1;
        if (o.first) o.first[p] = k;
    }
    if (k >= 0 && !accept_nearest(lg, no, k, d)) k = -1;
    int jT = -1, iT = -1;
    int8_t keep = 0;
    if (k >= 0) { jT = k / lg.Ni; iT = k - jT * lg.Ni; keep = 1; }
    if (o.nearest) o.nearest[p] = make_int2(jT, iT);
    if (o.dmin) o.dmin[p] = d;
    int cj = 0, ci = 0;
    if (keep && do_survive && killed(jT, iT, g.Nj, g.Ni, g.tmask, ic0, g.rmin_conc)) keep = 0;   // tracking.py:149
    if (keep && do_cell) {                                    // locate.py:280-330
        const pt yx = SC[p];
        const int dj[5] = {0, 0, 1, 0, -1}, di[5] = {0, 1, 0, -1, 0};
        bool in = false;
        for (int kp = 0; kp < 5 && !in; ++kp) {
            cj = jT + dj[kp]; ci = iT + di[kp];
            const int c = cj * g.Ni + ci;
            in = inside_quad(yx.y, yx.x, ldg_pt(g.F, c - g.Ni - 1), ldg_pt(g.F, c - g.Ni),
                             ldg_pt(g.F, c), ldg_pt(g.F, c - 1));
        }
        if (!in) keep = 0;                                    // tracking.py:156-160
    }
    if (o.cell) o.cell[p] = make_int2(cj, ci);
    if (o.keep) o.keep[p] = keep;
}

// k_nearest_brute: whole-grid scan, one block per buoy -- the reference's own
// O(nP*Nj*Ni) algorithm, kept as the on-device cross-check of the hash search.
__global__ void __launch_bounds__(256)
k_nearest_brute(const LocateGrid lg, long long nP, const pt* __restrict__ SG, int2* __restrict__ nearest,
                double* __restrict__ dmin)
{
    __shared__ double sd[256]; __shared__ int sk[256];
    const long long p = blockIdx.x;
    const pt ll = SG[p];
    const int n = lg.Nj * lg.Ni;
    double bd = INFINITY; int bk = 0x7fffffff;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const double d = haversine_km(ll.y, ll.x, lg.latT[k], lg.lonT[k]);
        if (d < bd || (d == bd && k < bk)) { bd = d; bk = k; }
    }
    sd[threadIdx.x] = bd; sk[threadIdx.x] = bk;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if (threadIdx.x < o) {
            const double d = sd[threadIdx.x + o]; const int k = sk[threadIdx.x + o];
            if (d < sd[threadIdx.x] || (d == sd[threadIdx.x] && k < sk[threadIdx.x])) { sd[threadIdx.x] = d; sk[threadIdx.x] = k; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        nearest[p] = make_int2(sk[0] / lg.Ni, sk[0] % lg.Ni);
        if (dmin) dmin[p] = sd[0];
    }
}

static NearestOpt g_default_opt = {2.5, 7};

cudaError_t launch_seed_locate_opt(const LocateGrid& lg, const AdvectGrid& g, const float* ic0,
                                   long long nP, const pt* SG, const pt* SC, const SeedOut& o,
                                   double rd_found_km, int max_itr, int do_survive, int do_cell, cudaStream_t st)
{
    if (nP <= 0) return cudaSuccess;
    NearestOpt no;
    no.rd_found_km = rd_found_km;
    // locate.py:250-275: radii r0, 1.2 r0, ... are tried at igo = 2..max_itr and a hit at
    // igo == max_itr is thrown away, so the last useful radius is r0 * 1.2^(max_itr-3).
    if (max_itr >= 3) no.accept_mode = max_itr - 3;
    else if (max_itr == 1) no.accept_mode = -2;
    else no.accept_mode = -1;
    double r0 = lg.resKM ? 0.5 * lg.res_max : rd_found_km;
    for (int q = 0; q < (no.accept_mode > 0 ? no.accept_mode : 0); ++q) r0 *= 1.2;
    const double r_accept_max = (no.accept_mode == -2) ? 1e30 : r0 * (1.0 + 1e-9);
    const int B = 128;
    k_seed_locate<<<(unsigned)((nP + B - 1) / B), B, 0, st>>>(lg, g, ic0, nP, SG, SC, o, no, r_accept_max,
                                                             do_survive, do_cell);
    return cudaGetLastError();
}

cudaError_t launch_seed_locate(const LocateGrid& lg, const AdvectGrid& g, const float* ic0,
                               long long nP, const pt* SG, const pt* SC, const SeedOut& o,
                               int do_survive, int do_cell, cudaStream_t st)
{
    // SeedInit's call: rd_found_km=2.5 (overridden by 0.5*resKM), max_itr=10 (tracking.py:134)
    return launch_seed_locate_opt(lg, g, ic0, nP, SG, SC, o, g_default_opt.rd_found_km, 10, do_survive, do_cell, st);
}

// SeedInit's shrink to the kept buoys (tracking.py:166-178) on the device: pos and cell compacted by `keep`, order kept
cudaError_t seed_compact(long long nP, const pt* pos, const int2* cell, const int8_t* keep, pt* out_pos, int2* out_cell,
                         long long* d_nout, cudaStream_t st)
{
    if (nP <= 0) return cudaMemsetAsync(d_nout, 0, sizeof(long long), st);
    void* tmp = nullptr;
    size_t b1 = 0, b2 = 0;
    const double2* p2 = reinterpret_cast<const double2*>(pos);
    double2* o2 = reinterpret_cast<double2*>(out_pos);
    cudaError_t e = cub::DeviceSelect::Flagged(nullptr, b1, p2, keep, o2, d_nout, nP, st);
    if (e == cudaSuccess) e = cub::DeviceSelect::Flagged(nullptr, b2, cell, keep, out_cell, d_nout, nP, st);
    if (e == cudaSuccess) e = cudaMalloc(&tmp, b1 > b2 ? b1 : b2);
    if (e == cudaSuccess) e = cub::DeviceSelect::Flagged(tmp, b1, p2, keep, o2, d_nout, nP, st);
    if (e == cudaSuccess) e = cub::DeviceSelect::Flagged(tmp, b2, cell, keep, out_cell, d_nout, nP, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(tmp);
    return e;
}

cudaError_t launch_nearest_brute(const LocateGrid& lg, long long nP, const pt* SG, int2* nearest,
                                 double* dmin, cudaStream_t st)
{
    if (nP <= 0) return cudaSuccess;
    k_nearest_brute<<<(unsigned)nP, 256, 0, st>>>(lg, nP, SG, nearest, dmin);
    return cudaGetLastError();
}

}  // namespace st
