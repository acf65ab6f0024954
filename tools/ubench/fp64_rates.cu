// Micro-benchmark: per-SM issue rates of DFMA, F2F.F64.F32, F2F.F32.F64 and an integer-emulated f32->f64
// conversion on sm_100a (to decide how the factor kernels should widen their float32 inputs).
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double* out, const float* in, int iters) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    double a0 = in[t & 255], a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    float f0 = in[(t + 1) & 255], f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3, f4 = f0 + 4, f5 = f0 + 5, f6 = f0 + 6, f7 = f0 + 7;
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {  // 8 independent DFMA chains
            a0 = fma(a0, 1.0000001, 1e-9); a1 = fma(a1, 1.0000001, 1e-9); a2 = fma(a2, 1.0000001, 1e-9); a3 = fma(a3, 1.0000001, 1e-9);
            a4 = fma(a4, 1.0000001, 1e-9); a5 = fma(a5, 1.0000001, 1e-9); a6 = fma(a6, 1.0000001, 1e-9); a7 = fma(a7, 1.0000001, 1e-9);
        } else if (MODE == 1) {  // f32 -> f64 conversions feeding cheap float updates
            a0 += (double)f0; f0 *= 1.0001f; a1 += (double)f1; f1 *= 1.0001f; a2 += (double)f2; f2 *= 1.0001f; a3 += (double)f3; f3 *= 1.0001f;
            a4 += (double)f4; f4 *= 1.0001f; a5 += (double)f5; f5 *= 1.0001f; a6 += (double)f6; f6 *= 1.0001f; a7 += (double)f7; f7 *= 1.0001f;
        } else if (MODE == 2) {  // same with the integer widening
#define WIDE(f) __hiloint2double((int)((((unsigned)__float_as_int(f) >> 3) & 0x0fffffffu) + 0x38000000u) | (__float_as_int(f) & 0x80000000), __float_as_int(f) << 29)
            a0 += WIDE(f0); f0 *= 1.0001f; a1 += WIDE(f1); f1 *= 1.0001f; a2 += WIDE(f2); f2 *= 1.0001f; a3 += WIDE(f3); f3 *= 1.0001f;
            a4 += WIDE(f4); f4 *= 1.0001f; a5 += WIDE(f5); f5 *= 1.0001f; a6 += WIDE(f6); f6 *= 1.0001f; a7 += WIDE(f7); f7 *= 1.0001f;
        } else if (MODE == 3) {  // DADD only (baseline for modes 1, 2)
            a0 += 1e-9; a1 += 1e-9; a2 += 1e-9; a3 += 1e-9; a4 += 1e-9; a5 += 1e-9; a6 += 1e-9; a7 += 1e-9;
            f0 *= 1.0001f; f1 *= 1.0001f; f2 *= 1.0001f; f3 *= 1.0001f; f4 *= 1.0001f; f5 *= 1.0001f; f6 *= 1.0001f; f7 *= 1.0001f;
        } else if (MODE == 4) {  // f64 -> f32
            f0 += (float)a0; a0 += 1e-9; f1 += (float)a1; a1 += 1e-9; f2 += (float)a2; a2 += 1e-9; f3 += (float)a3; a3 += 1e-9;
            f4 += (float)a4; a4 += 1e-9; f5 += (float)a5; a5 += 1e-9; f6 += (float)a6; a6 += 1e-9; f7 += (float)a7; a7 += 1e-9;
        } else if (MODE == 5) {  // one dependent DFMA chain (latency)
            a0 = fma(a0, 1.0000001, 1e-9); a0 = fma(a0, 1.0000001, 1e-9); a0 = fma(a0, 1.0000001, 1e-9); a0 = fma(a0, 1.0000001, 1e-9);
            a0 = fma(a0, 1.0000001, 1e-9); a0 = fma(a0, 1.0000001, 1e-9); a0 = fma(a0, 1.0000001, 1e-9); a0 = fma(a0, 1.0000001, 1e-9);
        } else if (MODE == 6) {  // two dependent chains
            a0 = fma(a0, 1.0000001, 1e-9); a1 = fma(a1, 1.0000001, 1e-9); a0 = fma(a0, 1.0000001, 1e-9); a1 = fma(a1, 1.0000001, 1e-9);
            a0 = fma(a0, 1.0000001, 1e-9); a1 = fma(a1, 1.0000001, 1e-9); a0 = fma(a0, 1.0000001, 1e-9); a1 = fma(a1, 1.0000001, 1e-9);
        }
    }
    out[t] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7;
}
template <int MODE>
void run(const char* name, int warps_per_sm, double* out, float* in) {
    const int iters = 20000, sms = 148;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms, warps_per_sm * 32>>>(out, in, 100);
    cudaEventRecord(e0);
    k<MODE><<<sms, warps_per_sm * 32>>>(out, in, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double cycles = ms * 1e-3 * 1.965e9;
    printf("%-28s warps/SM %2d: %8.3f ms, %7.2f cycles per loop iteration (8 ops/thread), %6.2f lane-ops/clk/SM\n", name,
           warps_per_sm, ms, cycles / iters, 8.0 * warps_per_sm * 32 * iters / cycles);
}
int main() {
    double* out; float* in;
    cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&in, 1024);
    cudaMemset(in, 0, 1024);
    for (int w : {4, 8, 16}) {
        if (w == 4) { run<0>("DFMA x8 indep", 4, out, in); run<1>("F2F.F64.F32 + DADD + FMUL", 4, out, in); run<2>("int-widen + DADD + FMUL", 4, out, in); run<3>("DADD + FMUL", 4, out, in); run<4>("F2F.F32.F64 + FADD + DADD", 4, out, in); run<5>("DFMA 1 chain", 4, out, in); run<6>("DFMA 2 chains", 4, out, in); }
        if (w == 8) { run<0>("DFMA x8 indep", 8, out, in); run<1>("F2F.F64.F32 + DADD + FMUL", 8, out, in); run<2>("int-widen + DADD + FMUL", 8, out, in); run<3>("DADD + FMUL", 8, out, in); run<4>("F2F.F32.F64 + FADD + DADD", 8, out, in); run<5>("DFMA 1 chain", 8, out, in); run<6>("DFMA 2 chains", 8, out, in); }
        if (w == 16) { run<0>("DFMA x8 indep", 16, out, in); run<1>("F2F.F64.F32 + DADD + FMUL", 16, out, in); run<2>("int-widen + DADD + FMUL", 16, out, in); run<3>("DADD + FMUL", 16, out, in); run<4>("F2F.F32.F64 + FADD + DADD", 16, out, in); run<5>("DFMA 1 chain", 16, out, in); run<6>("DFMA 2 chains", 16, out, in); }
    }
    return 0;
}
