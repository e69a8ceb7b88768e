// Memory-system skeleton of the step kernel: the same streams (read pos 16 B + cell 8 B; write pos 16 B, yx 16 B,
// latlon 16 B, mask 1 B per buoy), no gathers, no arithmetic.  What this achieves is the practical ceiling of the
// step's access mix on this GPU (the MEASURED_PEAKS copy figure is a 1:1 read/write copy of one stream).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o stream_skeleton tools/micro/stream_skeleton.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

// SP: sparse extra stores by ~13 % of the lanes (hash of the index):
//   1 lone 16-B .cs stores into a scratch array, 2 the same as whole 32-B sectors (lane pairs), 3 lone 8-B stores into the
//   cell array itself (what a cell update is), 4 lone 16-B plain stores, 5 = 3 but every lane of a marked 4-lane group
//   rewrites its cell (whole sectors)
template <int MODE, int SP = 0>   // 0: one thread per buoy, plain; 1: streaming hints (.cs); 2: persistent warp-per-tile with .cs
__global__ void k_skel(long long n, const double2* __restrict__ pos_in, int2* __restrict__ cell, double2* pos_out,
                       double2* yx, double2* ll, signed char* mask, int ntiles, double2* scratch = nullptr)
{
    if (MODE < 2) {
        const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
        if (p >= n) return;
        double2 P; int2 c;
        if (MODE == 0) { P = pos_in[p]; c = cell[p]; } else { P = __ldcs(pos_in + p); c = __ldcs(cell + p); }
        P.x += (double)c.x * 1e-9; P.y += (double)c.y * 1e-9;
        const double2 L = make_double2(P.x * 0.5, P.y * 0.25);
        if (MODE == 0) { pos_out[p] = P; yx[p] = P; ll[p] = L; mask[p] = 1; }
        else { __stcs(pos_out + p, P); __stcs(yx + p, P); __stcs(ll + p, L); __stcs(mask + p, (signed char)1); }
        if (SP) {
            const bool mk = ((unsigned)(p * 2654435761u) >> 16) % 100 < 13;
            if (SP == 1) { if (mk) __stcs(scratch + p, P); }
            if (SP == 4) { if (mk) scratch[p] = P; }
            if (SP == 2) { if (__shfl_xor_sync(0xffffffffu, (int)mk, 1) | (int)mk) __stcs(scratch + p, P); }
            if (SP == 3) { if (mk) __stcs(cell + p, make_int2(c.x + 1, c.y)); }
            if (SP == 5) { int g = (int)mk; g |= __shfl_xor_sync(0xffffffffu, g, 1); g |= __shfl_xor_sync(0xffffffffu, g, 2);
                           if (g) __stcs(cell + p, make_int2(c.x + (int)mk, c.y)); }
        }
    } else {
        const int lane = threadIdx.x & 31;
        const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
        for (int t = w; t < ntiles; t += nw) {
            const long long p = (long long)t * 32 + lane;
            if (p >= n) break;
            double2 P = __ldcs(pos_in + p); const int2 c = __ldcs(cell + p);
            P.x += (double)c.x * 1e-9; P.y += (double)c.y * 1e-9;
            const double2 L = make_double2(P.x * 0.5, P.y * 0.25);
            __stcs(pos_out + p, P); __stcs(yx + p, P); __stcs(ll + p, L); __stcs(mask + p, (signed char)1);
        }
    }
}

int main(int argc, char** argv)
{
    const long long n = argc > 1 ? atoll(argv[1]) : 12469235;
    double2 *pos, *yx, *ll; int2* cell; signed char* mask;
    CK(cudaMalloc(&pos, n * 16)); CK(cudaMalloc(&yx, n * 16)); CK(cudaMalloc(&ll, n * 16));
    CK(cudaMalloc(&cell, n * 8)); CK(cudaMalloc(&mask, n));
    CK(cudaMemset(pos, 0, n * 16)); CK(cudaMemset(cell, 0, n * 8));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int ntiles = (int)((n + 31) / 32);
    int nsm = 0; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
    double2* scratch; CK(cudaMalloc(&scratch, n * 16));
    for (int mode = 0; mode < 11; ++mode) {
        float best = 1e9f, sum = 0; const int reps = 30;
        for (int r = 0; r < reps + 3; ++r) {
            CK(cudaEventRecord(e0));
            if (mode == 0) k_skel<0><<<(unsigned)((n + 255) / 256), 256>>>(n, pos, cell, pos, yx, ll, mask, ntiles);
            else if (mode == 1) k_skel<1><<<(unsigned)((n + 255) / 256), 256>>>(n, pos, cell, pos, yx, ll, mask, ntiles);
            else if (mode == 2) k_skel<2><<<nsm * 32, 32>>>(n, pos, cell, pos, yx, ll, mask, ntiles);
            else if (mode == 3) k_skel<2><<<nsm * 24, 32>>>(n, pos, cell, pos, yx, ll, mask, ntiles);
            else if (mode == 4) k_skel<2><<<nsm * 16, 32>>>(n, pos, cell, pos, yx, ll, mask, ntiles);
            else if (mode == 5) k_skel<2><<<nsm * 8, 256>>>(n, pos, cell, pos, yx, ll, mask, ntiles);
            else if (mode == 6) k_skel<1, 1><<<(unsigned)((n + 255) / 256), 256>>>(n, pos, cell, pos, yx, ll, mask, ntiles, scratch);
            else if (mode == 7) k_skel<1, 2><<<(unsigned)((n + 255) / 256), 256>>>(n, pos, cell, pos, yx, ll, mask, ntiles, scratch);
            else if (mode == 8) k_skel<1, 3><<<(unsigned)((n + 255) / 256), 256>>>(n, pos, cell, pos, yx, ll, mask, ntiles, scratch);
            else if (mode == 9) k_skel<1, 4><<<(unsigned)((n + 255) / 256), 256>>>(n, pos, cell, pos, yx, ll, mask, ntiles, scratch);
            else k_skel<1, 5><<<(unsigned)((n + 255) / 256), 256>>>(n, pos, cell, pos, yx, ll, mask, ntiles, scratch);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (r >= 3) { sum += ms; if (ms < best) best = ms; }
        }
        const double bytes = 81.0 * n;
        printf("mode %d: mean %.1f us (%.0f GB/s)  best %.1f us (%.0f GB/s)\n", mode, sum / reps * 1e3,
               bytes / (sum / reps * 1e-3) / 1e9, best * 1e3, bytes / (best * 1e-3) / 1e9);
    }
    return 0;
}
