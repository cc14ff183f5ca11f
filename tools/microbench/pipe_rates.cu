// Issue-rate microbenchmark for the instructions the matcher epilogue is made of (sm_100a).
// Prints warp-instructions per clock per SM for each op with 4 warps per scheduler.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 4096, UNROLL = 8;

template <int OP>
__global__ void __launch_bounds__(512) bench(uint32_t *out, long long *cycles, uint32_t seed) {
    uint32_t a[UNROLL], b = seed ^ threadIdx.x, c = seed * 3 + threadIdx.x;
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) a[i] = seed + i * 77 + threadIdx.x * 13;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < UNROLL; ++i) {
            if (OP == 0) a[i] = max(a[i], b + it);                                  // VIMNMX (+ IADD outside? b+it hoisted per it)
            if (OP == 1) a[i] = max(max(a[i], b), c ^ a[(i + 1) % UNROLL]);          // VIMNMX3 + LOP3
            if (OP == 2) asm volatile("redux.sync.max.u32 %0, %0, 0xffffffff;" : "+r"(a[i]));
            if (OP == 3) asm volatile("shfl.sync.bfly.b32 %0, %0, 16, 0x1f, 0xffffffff;" : "+r"(a[i]));
            if (OP == 4) { uint32_t r; asm volatile("{.reg .pred p; setp.ne.u32 p, %1, 0; vote.sync.ballot.b32 %0, p, 0xffffffff;}" : "=r"(r) : "r"(a[i])); a[i] += r; }
            if (OP == 5) a[i] = (a[i] & 0xffffffe0u) | b;                            // LOP3
            if (OP == 6) a[i] = __float_as_uint(__uint_as_float(a[i]) + __uint_as_float(b));  // FADD
            if (OP == 7) a[i] = a[i] * 3 + b;                                        // IMAD
            if (OP == 8) a[i] = (__uint_as_float(a[i]) > __uint_as_float(c)) ? a[i] : b;      // FSETP + SEL
            if (OP == 9) { a[i] = max(a[i], b); a[i] = __float_as_uint(__uint_as_float(a[i]) + 1.0f); }  // VIMNMX + FADD alternating
            if (OP == 10) { uint32_t m = min(a[i], b); a[i] = max(a[i], c) + m; }    // 2 VIMNMX + IADD
            if (OP == 11) a[i] = max(max(a[i], b), c + i);                            // pure VIMNMX3
            if (OP == 12) { asm volatile("redux.sync.max.u32 %0, %0, 0xffffffff;" : "+r"(a[i])); a[i] = max(a[i] ^ b, c); a[i] = max(a[i] + 1, b); a[i] = min(a[i] + 3, c); a[i] = max(a[i] + 5, b);}  // 1 REDUX : 4 VIMNMX(+adds)
        }
        b += 0x9e3779b9u; c ^= b;
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < UNROLL; ++i) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char *name, double instr_per_iter) {
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
    bench<OP><<<148, 512>>>(out, cyc, 12345u);
    bench<OP><<<148, 512>>>(out, cyc, 12345u);
    cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    const double warp_instr = 16.0 * ITERS * UNROLL * instr_per_iter;   // 16 warps per SM
    printf("%-28s %8.0f cycles  %.3f listed-instr/clk/SM  (%.2f clk per warp-instr per SMSP)\n", name, avg, warp_instr / avg,
           avg / (warp_instr / 4));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("VIMNMX", 1); run<11>("VIMNMX3", 1); run<1>("VIMNMX3+LOP3", 2); run<2>("REDUX.MAX", 1); run<3>("SHFL.BFLY", 1);
    run<4>("VOTE.ballot(+setp+add)", 3); run<5>("LOP3", 1); run<6>("FADD", 1); run<7>("IMAD", 1); run<8>("FSETP+SEL", 2);
    run<9>("VIMNMX+FADD", 2); run<10>("2xVIMNMX+IADD", 3); run<12>("REDUX+4x(VIMNMX+LOP/ADD)", 9);
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
