// sweep_probe.cu -- measurement harness (not product): denominators and structural probes for the sweep kernels.
//   1. FP64 FMA issue rate per SM (the compute roofline of dense-gate rounds; SURVEY 8d asks for a measured figure)
//   2. warp-shuffle and LDS/STS.128 rates
//   3. both sweep kernels with EMPTY programs (no gate arithmetic) for tile-row sizes 512 B .. 64 KB and 1..4 rounds:
//      what the data movement alone costs
// build: make -C scripts/micro ; run on the GPU box: scripts/micro/sweep_probe [n]
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#include "sv_kernels.cuh"

using namespace b200;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { std::printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); std::exit(1); } } while (0)

template <int CHAINS>
__global__ void __launch_bounds__(1024) dfma_kernel(double* out, const int iters, const double a, const double b) {
    double x[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) x[k] = threadIdx.x * 1e-3 + k;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < CHAINS; ++k) x[k] = fma(x[k], a, b);
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) s += x[k];
    if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(1024) shfl_kernel(double* out, const int iters) {
    double x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = threadIdx.x + k;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = __shfl_xor_sync(0xffffffffu, x[k], 1 + (i & 3));
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += x[k];
    if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(256) lds_kernel(double* out, const int iters) {
    extern __shared__ __align__(16) double2 sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = make_double2(i, 0);
    __syncthreads();
    double2 acc = make_double2(0, 0);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const double2 v = sm[(threadIdx.x + 256 * k + i) & 4095];
            acc.x += v.x; acc.y += v.y;
        }
    }
    if (acc.x == 12345.678) out[0] = acc.x + acc.y;
}

static float time_ms(cudaStream_t st, const std::function<void()>& f, int reps = 3) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaStreamSynchronize(st);
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(a, st);
        f();
        cudaEventRecord(b, st);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(a); cudaEventDestroy(b);
    return best;
}

// program without ops: tile = c contiguous low qubits + the highest (12 - c) qubits; `nr` rounds
static SweepProg empty_prog(int n, int c, int nr) {
    SweepProg sp;
    std::memset(&sp, 0, sizeof sp);
    sp.nrounds = nr; sp.c = c;
    for (int i = 0; i < c; ++i) sp.tileq[i] = i;
    for (int i = c; i < TILE_BITS; ++i) sp.tileq[i] = n - (TILE_BITS - i);
    for (int r = 0; r < nr; ++r) {
        PRound& rd = sp.rounds[r];
        const bool hbm = r == 0 || r == nr - 1;
        for (int k = 0; k < REG_BITS; ++k) rd.regpos[k] = hbm ? TILE_BITS - REG_BITS + k : 2 * k + (r & 1);   // middle rounds: scattered low positions
        for (int j = 0; j < 16; ++j) {
            uint32_t off = 0;
            for (int b = 0; b < REG_BITS; ++b) if (j >> b & 1) off |= 1u << rd.regpos[b];
            rd.soff[j] = (int32_t)swz(off);
        }
    }
    return sp;
}

int main(int argc, char** argv) {
    const int n = argc > 1 ? std::atoi(argv[1]) : 28;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    std::printf("device %s, %d SMs, max clock %.0f MHz\n", prop.name, sms, clk_khz / 1e3);
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    double* d_out;
    CK(cudaMalloc(&d_out, 64));

    // ---- 1. FP64 FMA ----
    for (int threads : {256, 512, 1024}) {
        const int iters = 20000;
        const float ms = time_ms(st, [&] { dfma_kernel<16><<<sms, threads, 0, st>>>(d_out, iters, 1.0000001, 1e-9); });
        const double fma = (double)sms * threads * 16.0 * iters;
        std::printf("dfma  threads/SM=%4d  %.3f ms  %.2f TFLOP/s  (%.1f FMA/clk/SM at max clock)\n", threads, ms,
                    2 * fma / (ms * 1e-3) / 1e12, fma / (ms * 1e-3) / sms / (clk_khz * 1e3));
    }
    // ---- 2. shuffles, LDS ----
    {
        const int iters = 20000;
        const float ms = time_ms(st, [&] { shfl_kernel<<<sms, 1024, 0, st>>>(d_out, iters); });
        const double sh = (double)sms * 32 * 8 * 2.0 * iters;   // 32-bit warp shuffles (a double = 2)
        std::printf("shfl  %.3f ms  %.2f warp-SHFL.32/clk/SM\n", ms, sh / (ms * 1e-3) / sms / (clk_khz * 1e3));
        CK(cudaFuncSetAttribute(lds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        const int it2 = 4000;
        const float ms2 = time_ms(st, [&] { lds_kernel<<<sms * 2, 256, 65536, st>>>(d_out, it2); });
        const double bytes = (double)sms * 2 * 256 * 16.0 * it2 * 16;
        std::printf("lds.128  %.3f ms  %.1f B/clk/SM\n", ms2, bytes / (ms2 * 1e-3) / sms / (clk_khz * 1e3));
    }
    // ---- 3. sweep kernels, empty programs ----
    const uint64_t dim = 1ull << n;
    double2 *a, *b;
    CK(cudaMalloc(&a, dim * sizeof(double2)));
    CK(cudaMalloc(&b, dim * sizeof(double2)));
    CK(cudaMemset(a, 0, dim * sizeof(double2)));
    CK(cudaMemset(b, 0, dim * sizeof(double2)));
    CK(cudaFuncSetAttribute(sv_sweep_kernel<REG_BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TILE_BYTES));
    CK(cudaFuncSetAttribute(sv_sweep_pipe_kernel<REG_BITS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(PIPE_STAGES * TILE_BYTES)));
    const uint32_t ntiles = (uint32_t)(dim >> TILE_BITS);
    const float msc = time_ms(st, [&] { cudaMemcpyAsync(b, a, dim * sizeof(double2), cudaMemcpyDeviceToDevice, st); });
    std::printf("cudaMemcpy D2D  %.3f ms  %.0f GB/s\n", msc, 32.0 * dim / (msc * 1e-3) / 1e9);
    for (int c : {5, 6, 7, 8, 10, 12}) {
        for (int nr : {1, 2, 3, 4}) {
            const SweepProg sp = empty_prog(n, c, nr);
            const float md = time_ms(st, [&] {
                sv_sweep_kernel<REG_BITS><<<sms * 2, SWEEP_THREADS, nr > 1 ? TILE_BYTES : 0, st>>>(a, b, sp, ntiles, EmbedSrc{}); });
            const float mp = time_ms(st, [&] {
                sv_sweep_pipe_kernel<REG_BITS><<<sms, PIPE_THREADS, PIPE_STAGES * TILE_BYTES, st>>>(a, b, sp, ntiles); });
            std::printf("empty sweep  c=%2d (row %5d B) rounds=%d  direct %.3f ms %6.0f GB/s | pipe %.3f ms %6.0f GB/s\n", c, 16 << c, nr,
                        md, 32.0 * dim / (md * 1e-3) / 1e9, mp, 32.0 * dim / (mp * 1e-3) / 1e9);
            CK(cudaGetLastError());
        }
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
