// ubench.cu — B200 pipe-throughput probes used to size the FAST kernel's instruction budget.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench ubench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("%s: %s\n",#x,cudaGetErrorString(e)); exit(1);} }while(0)

constexpr int ITERS = 4096, ILP = 8;

template <int OP>
__global__ void __launch_bounds__(1024) k(float* out, float a, float b, long long* cyc) {
    float x[ILP]; float2 v[ILP];
    for (int i = 0; i < ILP; ++i) { x[i] = threadIdx.x * 0.001f + i; v[i] = make_float2(x[i], x[i] + 0.5f); }
    const float2 a2 = make_float2(a, a * 1.0001f), b2 = make_float2(b, b * 0.999f);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) {
            if (OP == 0) x[i] = __fmaf_rn(x[i], a, b);
            if (OP == 1) x[i] = __fadd_rn(x[i], a);
            if (OP == 2) x[i] = __fmul_rn(x[i], a);
            if (OP == 3) v[i] = __ffma2_rn(v[i], a2, b2);
            if (OP == 4) v[i] = __fadd2_rn(v[i], a2);
            if (OP == 5) v[i] = __fmul2_rn(v[i], a2);
            if (OP == 6) asm volatile("rsqrt.approx.f32 %0, %0;" : "+f"(x[i]));
            if (OP == 7) asm volatile("rcp.approx.f32 %0, %0;" : "+f"(x[i]));
            if (OP == 8) { x[i] = __fmaf_rn(x[i], a, b); v[i] = __fadd2_rn(v[i], a2); }   // mix
            if (OP == 9) { asm volatile("{.reg .pred p; setp.gt.f32 p, %0, %1; selp.f32 %0, %0, %2, p;}" : "+f"(x[i]) : "f"(a), "f"(b)); }
            if (OP == 10) { unsigned u = __float_as_uint(x[i]); u = u * 0xD2511F53u + 12345u; x[i] = __uint_as_float(u); }
            if (OP == 11) { unsigned u = __float_as_uint(x[i]); u = __umulhi(u, 0xD2511F53u) ^ (u * 0xCD9E8D57u); x[i] = __uint_as_float(u); }
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < ILP; ++i) s += x[i] + v[i].x + v[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

// streaming 16 B/cell: read 2 arrays, write 2 arrays, float4
__global__ void stream4(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ c, float4* __restrict__ d, long long n) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        float4 x = a[i], y = b[i];
        c[i] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
        d[i] = make_float4(x.x - y.x, x.y - y.y, x.z - y.z, x.w - y.w);
    }
}

int main() {
    int dev = 0; cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
    printf("device %s SMs %d clock %d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    float* out; long long* cyc; CK(cudaMalloc(&out, sizeof(float) * 148 * 8 * 1024)); CK(cudaMalloc(&cyc, 8));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const char* names[] = {"FFMA", "FADD", "FMUL", "FFMA2", "FADD2", "FMUL2", "MUFU.RSQ", "MUFU.RCP", "FFMA+FADD2", "FSETP+SEL", "IMAD", "IMAD.HI+IMAD+LOP"};
    for (int op = 0; op < 12; ++op) {
        const int blocks = p.multiProcessorCount * 2, threads = 1024;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            switch (op) {
#define C(n) case n: k<n><<<blocks, threads>>>(out, 1.0001f, 0.0001f, cyc); break;
                C(0) C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11)
            }
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
        double warp_instr_per_sm = 2.0 * 32 * ITERS * ILP * ((op == 8 || op==9) ? 2 : (op == 11 ? 3 : 1));  // 2 CTAs x 32 warps
        printf("%-18s %8.3f ms  cycles(1 CTA) %lld  -> %.2f warp-instr/clk/SM (by clock64), MHz est %.0f\n", names[op], ms, c,
               warp_instr_per_sm / (double)c, (double)c / (ms * 1e3));
    }
    // streaming
    const long long n4 = (1LL << 28) / 4;  // 2^28 floats per array = 1 GiB
    float4 *a, *b, *c, *d;
    CK(cudaMalloc(&a, n4 * 16)); CK(cudaMalloc(&b, n4 * 16)); CK(cudaMalloc(&c, n4 * 16)); CK(cudaMalloc(&d, n4 * 16));
    CK(cudaMemset(a, 0, n4 * 16)); CK(cudaMemset(b, 0, n4 * 16));
    for (int blocks : {148 * 4, 148 * 8, 148 * 16, 148 * 32}) {
        for (int threads : {256, 512}) {
            float best = 1e9;
            for (int rep = 0; rep < 5; ++rep) {
                cudaEventRecord(e0); stream4<<<blocks, threads>>>(a, b, c, d, n4); cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            printf("stream4 blocks %5d threads %4d: %.3f ms  %.1f GB/s  (%.1f Gcell/s at 16 B/cell)\n", blocks, threads, best,
                   4.0 * n4 * 16 / best / 1e6, n4 * 4 / best / 1e6);
        }
    }
    return 0;
}
