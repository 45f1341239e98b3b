// peak.cu — FP32 FMA-pipe peak microbenchmark (roofline denominator for K1; BASELINE.md §2 asks
// the builder to measure it because MEASURED_PEAKS.json has no FP32 entry).
// variant 0: scalar FFMA chains; 1: packed fma.rn.f32x2 (FFMA2); 2: mul.f32x2; 3: add.f32x2 (ops counted as 2 flop each).
#include "common.cuh"

namespace pnbx {
namespace {
constexpr int CHAINS = 8;

template <int VARIANT>
__global__ void __launch_bounds__(256) fma_chain(float* out, int iters, float a, float b) {
    float x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = (float)(threadIdx.x + c) * 1e-3f;
    if (VARIANT == 0) {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) x[c] = fmaf(x[c], a, b);
        }
    } else {
        unsigned long long ab, bb;
        asm("mov.b64 %0, {%1, %1};" : "=l"(ab) : "f"(a));
        asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
        unsigned long long p[CHAINS / 2];
#pragma unroll
        for (int c = 0; c < CHAINS / 2; ++c) asm("mov.b64 %0, {%1, %2};" : "=l"(p[c]) : "f"(x[2 * c]), "f"(x[2 * c + 1]));
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int c = 0; c < CHAINS / 2; ++c) {
                if (VARIANT == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[c]) : "l"(ab), "l"(bb));
                if (VARIANT == 2) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[c]) : "l"(ab));
                if (VARIANT == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[c]) : "l"(bb));
            }
        }
#pragma unroll
        for (int c = 0; c < CHAINS / 2; ++c) asm("mov.b64 {%0, %1}, %2;" : "=f"(x[2 * c]), "=f"(x[2 * c + 1]) : "l"(p[c]));
    }
    float s = 0.f;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    if (s == 12345.678f) out[0] = s;  // never true; keeps the chains alive
}
}  // namespace
}  // namespace pnbx

// Returns achieved TFLOP/s (2 flop per FMA) of the chosen variant on `device`.
extern "C" int pnbx_measure_fp32_peak(int device, int variant, double* tflops) {
    using namespace pnbx;
    return guarded([&] {
        pnbx_opts o{device, PNBX_MEM_DEVICE, 0, 0, nullptr};  // own stream, no copies
        Exec ex = make_exec(&o);
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ex.device);
        DevBuf<float> out(4, ex.stream);
        const int iters = 1 << 16;
        dim3 grid(sms * 8);
        cudaEvent_t a, b;
        PNBX_CUDA(cudaEventCreate(&a));
        PNBX_CUDA(cudaEventCreate(&b));
        double best = 0.0;
        for (int rep = 0; rep < 5; ++rep) {
            PNBX_CUDA(cudaEventRecord(a, ex.stream));
            if (variant == 0) fma_chain<0><<<grid, 256, 0, ex.stream>>>(out.get(), iters, 0.999f, 1e-3f);
            else if (variant == 1) fma_chain<1><<<grid, 256, 0, ex.stream>>>(out.get(), iters, 0.999f, 1e-3f);
            else if (variant == 2) fma_chain<2><<<grid, 256, 0, ex.stream>>>(out.get(), iters, 0.999f, 1e-3f);
            else fma_chain<3><<<grid, 256, 0, ex.stream>>>(out.get(), iters, 0.999f, 1e-3f);
            PNBX_CUDA(cudaEventRecord(b, ex.stream));
            PNBX_CUDA(cudaEventSynchronize(b));
            float ms = 0.f;
            PNBX_CUDA(cudaEventElapsedTime(&ms, a, b));
            double fl = 2.0 * CHAINS * (double)iters * 256.0 * grid.x;
            best = std::max(best, fl / (ms * 1e-3) / 1e12);
        }
        cudaEventDestroy(a);
        cudaEventDestroy(b);
        *tflops = best;
        PNBX_CUDA(cudaStreamSynchronize(ex.stream));
    });
}
