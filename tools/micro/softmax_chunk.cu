// Microbenchmark: the per-chunk arithmetic of the attention softmax (32 scores -> 16 packed bf16 pairs + row sum) on register
// data, with 1/2/4 warps per SM sub-partition, to separate the arithmetic floor from the kernel's synchronisation costs.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o softmax_chunk softmax_chunk.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) { __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&v); }
__device__ __forceinline__ float2 ex2_poly2(float2 x)
{
    const float2 magic = make_float2(12582912.0f, 12582912.0f);
    x.x = fmaxf(x.x, -125.0f); x.y = fmaxf(x.y, -125.0f);
    const float2 t = __fadd2_rn(x, magic);
    const float2 nn = __fadd2_rn(magic, make_float2(-t.x, -t.y));
    const float2 f = __fadd2_rn(x, nn);
    float2 p = make_float2(0x1.c3f76p-5f, 0x1.c3f76p-5f);
    p = __ffma2_rn(p, f, make_float2(0x1.f0de1ap-3f, 0x1.f0de1ap-3f));
    p = __ffma2_rn(p, f, make_float2(0x1.62f31ap-1f, 0x1.62f31ap-1f));
    p = __ffma2_rn(p, f, make_float2(0x1.fff692p-1f, 0x1.fff692p-1f));
    float2 r;
    r.x = __uint_as_float(__float_as_uint(t.x) * 0x800000u + __float_as_uint(p.x));
    r.y = __uint_as_float(__float_as_uint(t.y) * 0x800000u + __float_as_uint(p.y));
    return r;
}
template <int POLY>
__device__ __forceinline__ float prob_chunk(const uint32_t (&v)[32], uint32_t (&pk)[16], float c, float m_use)
{
    const float2 c2 = make_float2(c, c), nm2 = make_float2(-m_use, -m_use);
    float2 acc0 = make_float2(0.0f, 0.0f), acc1 = make_float2(0.0f, 0.0f);
#pragma unroll
    for (int q = 0; q < 32; q += 2) {
        const float2 x = __ffma2_rn(make_float2(__uint_as_float(v[q]), __uint_as_float(v[q + 1])), c2, nm2);
        float2 e;
        if (((q >> 1) & 3) < POLY) e = ex2_poly2(x);
        else e = make_float2(ex2(x.x), ex2(x.y));
        pk[q >> 1] = pack_bf16(e.x, e.y);
        if (q & 2) acc1 = __fadd2_rn(acc1, e);
        else acc0 = __fadd2_rn(acc0, e);
    }
    return (acc0.x + acc0.y) + (acc1.x + acc1.y);
}
template <int POLY>
__global__ void k(float* out, long long* cyc, int iters)
{
    uint32_t v[32], pk[16];
    for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(-0.01f * (threadIdx.x + i));
    float l = 0.f, m = 1.0f;
    uint32_t sink = 0;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        l += prob_chunk<POLY>(v, pk, 1.25f, m);
#pragma unroll
        for (int i = 0; i < 16; ++i) { sink ^= pk[i]; v[2 * i] += pk[i] & 1u; }       // keep everything live; v changes every iteration
    }
    long long t1 = clock64();
    out[threadIdx.x] = l + sink;
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main()
{
    float* out; long long* cyc; cudaMalloc(&out, 4096 * 4); cudaMallocManaged(&cyc, 8);
    const int iters = 4096;
    for (int poly = 0; poly < 3; ++poly) {
        printf("POLY=%d:", poly);
        for (int warps : {4, 8, 16}) {
            void (*fn)(float*, long long*, int) = poly == 0 ? k<0> : poly == 1 ? k<1> : k<2>;
            fn<<<1, warps * 32>>>(out, cyc, iters); cudaDeviceSynchronize();
            fn<<<1, warps * 32>>>(out, cyc, iters); cudaDeviceSynchronize();
            printf("  %dw/smsp: %7.1f cycles per 32-score chunk per warp (%.2f per sub-partition per score)", warps / 4,
                   (double)*cyc / iters, (double)*cyc / iters / 32.0 / (warps / 4));
        }
        printf("\n");
    }
    return 0;
}
