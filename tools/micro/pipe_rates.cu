// Pipe-rate microbenchmark for sm_100a: cycles per warp-instruction of MUFU.EX2, FFMA, FFMA2, FADD2, F2FP, FMNMX, IMAD
// with 1/2/4/8 warps per SM sub-partition (one CTA on one SM).  nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
constexpr int ITERS = 2048;
template <int OP> __global__ void k(float* out, long long* cyc) {
    float a[8], b = threadIdx.x * 1e-3f + 1.0f;
    float2 a2[8];
    uint32_t u[8];
    for (int i = 0; i < 8; ++i) { a[i] = b + i; a2[i] = make_float2(b + i, b - i); u[i] = threadIdx.x + i; }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 1) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(b));
            if (OP == 2) { unsigned long long& r = *reinterpret_cast<unsigned long long*>(&a2[i]); unsigned long long bb = *reinterpret_cast<unsigned long long*>(&a2[(i + 1) & 7]); asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(r) : "l"(bb)); }
            if (OP == 3) { unsigned long long& r = *reinterpret_cast<unsigned long long*>(&a2[i]); unsigned long long bb = *reinterpret_cast<unsigned long long*>(&a2[(i + 1) & 7]); asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(r) : "l"(bb)); }
            if (OP == 4) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "+r"(u[i]) : "f"(a[i]), "f"(b));
            if (OP == 5) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
            if (OP == 6) asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
            if (OP == 7) asm volatile("lg2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) s += a[i] + a2[i].x + a2[i].y + u[i];
    out[threadIdx.x] = s;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 4096 * 4); cudaMallocManaged(&cyc, 8);
    const char* names[] = {"MUFU.EX2", "FFMA", "FFMA2", "FADD2", "F2FP.BF16x2", "FMNMX", "IMAD", "MUFU.LG2"};
    for (int op = 0; op < 8; ++op) {
        printf("%-12s", names[op]);
        for (int warps : {4, 8, 16, 32}) {                        // 1, 2, 4, 8 warps per sub-partition
            void (*fn)(float*, long long*) = op == 0 ? k<0> : op == 1 ? k<1> : op == 2 ? k<2> : op == 3 ? k<3> : op == 4 ? k<4> : op == 5 ? k<5> : op == 6 ? k<6> : k<7>;
            fn<<<1, warps * 32>>>(out, cyc); cudaDeviceSynchronize();
            fn<<<1, warps * 32>>>(out, cyc); cudaDeviceSynchronize();
            double per_smsp = (double)*cyc / (ITERS * 8.0 * (warps / 4));   // cycles per warp-instruction per sub-partition
            printf("  %2dw/smsp: %6.2f cyc/inst", warps / 4, per_smsp);
        }
        printf("\n");
    }
    return 0;
}
