/* pipe_bench.cu — which sm_100a pipe executes what, and how fast (development tool, not part of the product).
 * Every kernel runs NCH independent dependency chains per thread of one instruction (or an alternating mix of two /
 * three) and reports warp instructions per cycle per SM with 8 warps per scheduler resident. Used to balance the
 * wide-node test of rt_traverse.h between the FMA and the ALU pipe (PRMT vs IDP.4A vs I2F byte extraction).
 *   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_bench pipe_bench.cu && ./pipe_bench */
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define NCH 8
#define ITERS 4096

enum Op { PRMT, IDP, FFMA, FFMA_IMM, FMNMX, FMNMX3, LOP3, IADD3, SEL, I2F, IMAD, FADD, FMUL, FSETP_OR, FFMA2, SHF, LEA_, VIMNMX };

template <int OP>
__device__ __forceinline__ void op(uint32_t &r, uint32_t &r2, uint32_t c, uint32_t d) {
    if (OP == PRMT) asm volatile("prmt.b32 %0, %0, %1, 0x7604;" : "+r"(r) : "r"(c));
    if (OP == IDP) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(r) : "r"(c), "r"(d));
    if (OP == FFMA) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(*(float *)&r) : "f"(__uint_as_float(c)), "f"(__uint_as_float(d)));
    if (OP == FFMA_IMM) asm volatile("fma.rn.f32 %0, %0, 0f3F800001, %1;" : "+f"(*(float *)&r) : "f"(__uint_as_float(d)));
    if (OP == FMNMX) asm volatile("min.f32 %0, %0, %1;" : "+f"(*(float *)&r) : "f"(__uint_as_float(c)));
    if (OP == FMNMX3) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(*(float *)&r) : "f"(__uint_as_float(c)), "f"(__uint_as_float(d)));
    if (OP == LOP3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r) : "r"(c), "r"(d));
    if (OP == IADD3) asm volatile("add.u32 %0, %0, %1;" : "+r"(r) : "r"(c));
    if (OP == SEL) asm volatile("{.reg .pred p; setp.ne.u32 p, %2, 0; selp.b32 %0, %0, %1, p;}" : "+r"(r) : "r"(c), "r"(d));
    if (OP == I2F) asm volatile("{.reg .u8 b; .reg .u16 h; .reg .u32 t; bfe.u32 t, %0, 8, 8; cvt.u16.u32 h, t; cvt.rn.f32.u16 %0, h;}" : "+r"(r));
    if (OP == IMAD) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r) : "r"(c), "r"(d));
    if (OP == FADD) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(*(float *)&r) : "f"(__uint_as_float(c)));
    if (OP == FMUL) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(*(float *)&r) : "f"(__uint_as_float(c)));
    if (OP == FSETP_OR) asm volatile("{.reg .pred p; setp.le.f32 p, %1, %2; @p or.b32 %0, %0, 4;}" : "+r"(r) : "f"(__uint_as_float(c)), "f"(__uint_as_float(d)));
    if (OP == FFMA2) {
        uint64_t v = ((uint64_t)r2 << 32) | r, cc = ((uint64_t)c << 32) | c, dd = ((uint64_t)d << 32) | d;
        asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(cc), "l"(dd));
        r = (uint32_t)v;
        r2 = (uint32_t)(v >> 32);
    }
    if (OP == SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(r) : "r"(c), "r"(d));
    if (OP == VIMNMX) asm volatile("min.u32 %0, %0, %1;" : "+r"(r) : "r"(c));
}

template <int A, int B, int C>
__global__ void __launch_bounds__(128, 8) k(uint32_t *out, uint32_t c, uint32_t d, int iters) {
    uint32_t r[NCH], r2[NCH], s[NCH], s2[NCH], t[NCH], t2[NCH];
#pragma unroll
    for (int i = 0; i < NCH; i++) {
        r[i] = threadIdx.x * 7 + i + 0x3F800000u;
        s[i] = threadIdx.x * 5 + i + 0x3F800000u;
        t[i] = threadIdx.x * 3 + i + 0x3F800000u;
        r2[i] = s2[i] = t2[i] = 0x3F800000u;
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NCH; i++) {
            op<A>(r[i], r2[i], c, d);
            if (B >= 0) op<(B >= 0 ? B : 0)>(s[i], s2[i], c, d);
            if (C >= 0) op<(C >= 0 ? C : 0)>(t[i], t2[i], c, d);
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int i = 0; i < NCH; i++) acc ^= r[i] ^ s[i] ^ t[i] ^ r2[i] ^ s2[i] ^ t2[i];
    if (acc == 0x12345u) out[0] = acc;
}

template <int A, int B, int C>
void run(const char *name, int n_ops, int sms, double ghz) {
    uint32_t *out;
    cudaMalloc(&out, 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<A, B, C><<<sms * 8, 128>>>(out, 0x3F800123u, 0x3F800456u, 64);
    cudaEventRecord(e0);
    k<A, B, C><<<sms * 8, 128>>>(out, 0x3F800123u, 0x3F800456u, ITERS);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_insts = (double)sms * 8 * 4 * ITERS * NCH * n_ops;
    const double cycles = ms * 1e-3 * ghz * 1e9;
    printf("%-28s %8.3f ms  %6.3f warp-inst/clk/SM  (%5.3f per scheduler)\n", name, ms, warp_insts / cycles / sms, warp_insts / cycles / sms / 4);
    cudaFree(out);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz * 1e-6;
    printf("%s, %d SMs, %.3f GHz (attribute; the run is not clock-locked)\n", p.name, p.multiProcessorCount, ghz);
    const int sms = p.multiProcessorCount;
#define R1(A) run<A, -1, -1>(#A, 1, sms, ghz)
#define R2(A, B) run<A, B, -1>(#A "+" #B, 2, sms, ghz)
#define R3(A, B, C) run<A, B, C>(#A "+" #B "+" #C, 3, sms, ghz)
    R1(PRMT); R1(IDP); R1(FFMA); R1(FFMA_IMM); R1(FFMA2); R1(FMNMX); R1(FMNMX3); R1(LOP3); R1(IADD3); R1(SEL); R1(I2F); R1(IMAD);
    R1(FADD); R1(FMUL); R1(FSETP_OR); R1(SHF); R1(VIMNMX);
    R2(PRMT, FFMA); R2(IDP, FFMA); R2(PRMT, IDP); R2(PRMT, FMNMX); R2(IDP, FMNMX); R2(FFMA, FMNMX); R2(FFMA2, FMNMX); R2(FFMA2, PRMT);
    R2(I2F, FFMA); R2(I2F, PRMT); R2(IMAD, FFMA); R2(IDP, IMAD); R2(FADD, FMNMX); R2(FSETP_OR, FFMA);
    R3(PRMT, FFMA, FMNMX); R3(IDP, FFMA, FMNMX); R3(IDP, FFMA2, FMNMX); R3(I2F, FFMA, FMNMX); R3(PRMT, IDP, FMNMX);
    return 0;
}
