// FP64 FMA throughput probe: the compute-roofline denominator for the assembly kernel (MEASURED_PEAKS.json carries HBM and
// bf16 tensor peaks only; the DG assembly is plain FP64 FMA work).
#include "common.cuh"

namespace rb {

__global__ void __launch_bounds__(256) fma_probe_kernel(int iters, double* __restrict__ out) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

// FP64 tensor-core (DMMA) throughput probe: mma.sync m8n8k4 f64 chains, 8 independent 8x8 accumulator tiles per warp.  Settles
// the "tensor cores only if they pay" clause with a number: the DG contractions are 6x6 blocks (an 8x8 tile is 56 % full) with
// structural zeros the scalar kernels skip, so DMMA pays only if its rate is several times the DFMA rate of fma_probe_kernel.
__global__ void __launch_bounds__(256) dmma_probe_kernel(int iters, double* __restrict__ out) {
    double c[8][2];
#pragma unroll
    for (int t = 0; t < 8; ++t) c[t][0] = c[t][1] = 0.0;
    const double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * (1 + (threadIdx.x & 7));
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int t = 0; t < 8; ++t)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[t][0]), "+d"(c[t][1])
                         : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int t = 0; t < 8; ++t) s += c[t][0] + c[t][1];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace rb

extern "C" int reyna_b200_dmma_probe(int iters, double* scratch_dev, double* flops, void* stream) {
    RB_REQUIRE(scratch_dev && flops && iters > 0, RB_E_ARG, "reyna_b200_dmma_probe: bad argument");
    const int blocks = rb::kNumSMs * 8, threads = 256;   // scratch must hold blocks*threads doubles
    rb::dmma_probe_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, scratch_dev);
    RB_LAUNCH_CHECK("dmma_probe_kernel");
    *flops = 2.0 * 8.0 * 8.0 * 4.0 * 8.0 * (double)iters * blocks * (threads / 32);   // 8 tiles of m8n8k4 per warp and iteration
    return 0;
}

extern "C" int reyna_b200_fma_probe(int iters, double* scratch_dev, double* flops, void* stream) {
    RB_REQUIRE(scratch_dev && flops && iters > 0, RB_E_ARG, "reyna_b200_fma_probe: bad argument");
    const int blocks = rb::kNumSMs * 8, threads = 256;   // scratch must hold blocks*threads doubles
    rb::fma_probe_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, scratch_dev);
    RB_LAUNCH_CHECK("fma_probe_kernel");
    *flops = 2.0 * 8.0 * (double)iters * blocks * threads;
    return 0;
}
