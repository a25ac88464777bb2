// rbphd_microbench.cu -- FP64 pipe microbenchmark exported by librbphd.so (instrumentation; not on the frame path).
// The per-frame update is FP64 CUDA-core arithmetic compiled WITHOUT contraction (-fmad=false), so the peak that
// bounds it is the unfused DMUL + DADD issue rate, not the DFMA figure of the data sheet.  bench.py runs this once
// per process and reports the kernel's FP64 instruction rate against it (roofline.fp64).
#include "../../include/rbphd.h"

#include <cuda_runtime.h>

namespace {

constexpr int kAcc = 8;        // independent dependency chains per thread
constexpr int kInner = 64;     // unrolled operations per chain per outer iteration

// mode 0: fused multiply-add chains; mode 1: multiply then add (two instructions, rounded separately);
// mode 2: add only
template <int MODE>
__global__ void __launch_bounds__(256) k_fp64(double* out, int outer, double a, double b)
{
    double acc[kAcc];
#pragma unroll
    for (int k = 0; k < kAcc; k++) acc[k] = (double)(threadIdx.x + k) * 1e-3;
    for (int it = 0; it < outer; it++) {
#pragma unroll
        for (int u = 0; u < kInner; u++) {
#pragma unroll
            for (int k = 0; k < kAcc; k++) {
                if (MODE == 0) acc[k] = __fma_rn(acc[k], a, b);
                else if (MODE == 1) acc[k] = __dadd_rn(__dmul_rn(acc[k], a), b);
                else acc[k] = __dadd_rn(acc[k], b);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < kAcc; k++) s += acc[k];
    if (s == 12345.678) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // keeps the chains alive
}

template <int MODE>
int run(int outer, int grid, double* dout, cudaStream_t st, double* ms_best)
{
    cudaEvent_t e0, e1;
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return 1;
    double best = 1e30;
    for (int rep = 0; rep < 6; rep++) {   // first repetitions warm up
        cudaEventRecord(e0, st);
        k_fp64<MODE><<<grid, 256, 0, st>>>(dout, outer, 0.999999, 1e-9);
        cudaEventRecord(e1, st);
        if (cudaEventSynchronize(e1) != cudaSuccess) return 1;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep >= 2 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ms_best = best;
    return 0;
}

}  // namespace

extern "C" int rbphd_bench_fp64(int device, int outer, double out6[6])
{
    if (!out6 || outer < 1) return RBPHD_ERR_ARGUMENT;
    if (cudaSetDevice(device) != cudaSuccess) return RBPHD_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return RBPHD_ERR_CUDA;
    const int grid = prop.multiProcessorCount * 8;   // 2048 threads per SM
    double* dout = nullptr;
    if (cudaMalloc(&dout, sizeof(double) * 256 * (size_t)grid) != cudaSuccess) return RBPHD_ERR_CUDA;
    cudaStream_t st;
    cudaStreamCreate(&st);
    double ms[3] = {0, 0, 0};
    int rc = run<0>(outer, grid, dout, st, &ms[0]) | run<1>(outer, grid, dout, st, &ms[1]) | run<2>(outer, grid, dout, st, &ms[2]);
    cudaStreamDestroy(st);
    cudaFree(dout);
    if (rc) return RBPHD_ERR_CUDA;
    const double ops = (double)grid * 256.0 * outer * kInner * kAcc;   // chain steps executed
    out6[0] = 2.0 * ops / (ms[0] * 1e-3) / 1e12;   // DFMA: TFLOP/s (2 flops per instruction)
    out6[1] = 2.0 * ops / (ms[1] * 1e-3) / 1e12;   // unfused DMUL + DADD: TFLOP/s (2 instructions, 2 flops)
    out6[2] = ops / (ms[2] * 1e-3) / 1e12;         // DADD alone: TFLOP/s
    out6[3] = ops / (ms[0] * 1e-3) / 1e12;         // FP64 thread-instructions per second (1e12), fused
    out6[4] = 2.0 * ops / (ms[1] * 1e-3) / 1e12;   // FP64 thread-instructions per second (1e12), unfused
    out6[5] = (double)prop.multiProcessorCount;
    return RBPHD_OK;
}
