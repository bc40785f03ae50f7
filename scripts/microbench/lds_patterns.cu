// Shared-memory wavefront micro-benchmark: how many cycles does one LDS.128 / LDS.64 / LDS.32 cost per warp
// for different lane -> address patterns (broadcast groups adjacent or interleaved)?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_patterns lds_patterns.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

template <int WORDS>   // 4 = LDS.128, 2 = LDS.64, 1 = LDS.32
__global__ void k(const int* __restrict__ lane_off, float* out, long long* cyc, int iters) {
    extern __shared__ __align__(16) float sm[];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = (float)i;
    __syncthreads();
    const int off = lane_off[threadIdx.x & 31];       // in floats, multiple of WORDS
    float acc = 0.f;
    const unsigned base = (unsigned)__cvta_generic_to_shared(sm + off);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const unsigned a = base + (unsigned)(u & 7) * 2048u;   // 8 different 2 KB windows, same bank pattern
            if (WORDS == 4) { float4 v; asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a)); acc += v.x + v.y + v.z + v.w; }
            else if (WORDS == 2) { float2 v; asm volatile("ld.volatile.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a)); acc += v.x + v.y; }
            else { float v; asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(v) : "r"(a)); acc += v; }
        }
    }
    long long t1 = clock64();
    if (acc == 12345.678f) out[0] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int WORDS>
double run(const int* h_off, const char* name, int sms) {
    int* d_off; float* d_out; long long* d_cyc;
    cudaMalloc(&d_off, 128); cudaMalloc(&d_out, 4); cudaMalloc(&d_cyc, 8 * sms);
    cudaMemcpy(d_off, h_off, 128, cudaMemcpyHostToDevice);
    const int iters = 2000, threads = 512;
    k<WORDS><<<sms, threads, 32768>>>(d_off, d_out, d_cyc, 10);
    k<WORDS><<<sms, threads, 32768>>>(d_off, d_out, d_cyc, iters);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
    const double per = (double)c / ((double)iters * 8 * (threads / 32));     // SM cycles per warp-level load
    printf("%-58s LDS.%-3d %6.2f cycles/warp-instr\n", name, WORDS * 32, per);
    cudaFree(d_off); cudaFree(d_out); cudaFree(d_cyc);
    return per;
}

int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    const int sms = pr.multiProcessorCount;
    int o[32];
    auto all = [&](const char* name) { run<4>(o, name, sms); };
    for (int l = 0; l < 32; ++l) o[l] = 0;                       all("all lanes same address");
    for (int l = 0; l < 32; ++l) o[l] = l * 4;                   all("32 distinct contiguous 16B");
    for (int l = 0; l < 32; ++l) o[l] = (l / 4) * 4;             all("8 distinct, groups of 4 ADJACENT lanes, conflict-free");
    for (int l = 0; l < 32; ++l) o[l] = (l % 8) * 4;             all("8 distinct, lane%8 (interleaved), conflict-free");
    for (int l = 0; l < 32; ++l) o[l] = (l / 8) * 4;             all("4 distinct, groups of 8 ADJACENT lanes");
    for (int l = 0; l < 32; ++l) o[l] = (l % 4) * 4;             all("4 distinct, lane%4 (interleaved)");
    for (int l = 0; l < 32; ++l) o[l] = (l % 4) * 260;           all("4 distinct, lane%4, stride 260 floats (kernel operand)");
    for (int l = 0; l < 32; ++l) o[l] = (l / 8) * 260;           all("4 distinct, lane/8, stride 260 floats");
    for (int l = 0; l < 32; ++l) o[l] = (l / 4) * 4 + ((l / 4) & 1) * 32 * 3;  all("8 distinct groups of 4, two rows (no conflict)");
    for (int l = 0; l < 32; ++l) o[l] = ((l / 4) % 4) * 4 + ((l / 16)) * 32;   all("8 distinct groups of 4, 2-way bank conflict");
    for (int l = 0; l < 32; ++l) o[l] = (l / 2) * 4;             all("16 distinct, pairs of adjacent lanes");
    for (int l = 0; l < 32; ++l) o[l] = (l % 16) * 4;            all("16 distinct, lane%16");
    for (int l = 0; l < 32; ++l) o[l] = (l % 4) * 2;  run<2>(o, "LDS.64: 4 distinct lane%4", sms);
    for (int l = 0; l < 32; ++l) o[l] = (l / 8) * 2;  run<2>(o, "LDS.64: 4 distinct lane/8", sms);
    for (int l = 0; l < 32; ++l) o[l] = l * 2;        run<2>(o, "LDS.64: 32 distinct contiguous", sms);
    for (int l = 0; l < 32; ++l) o[l] = (l % 4);      run<1>(o, "LDS.32: 4 distinct lane%4", sms);
    for (int l = 0; l < 32; ++l) o[l] = l;            run<1>(o, "LDS.32: 32 distinct contiguous", sms);
    return 0;
}
