// Issue-rate microbenchmarks for candidate inner loops of the tile kernel (sm_100a).
// Each variant compares NR "read words" held in registers against one structure word fetched from
// shared memory per iteration and accumulates "any match"; prints cells/s (lane-level compares/s).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench tools/microbench.cu && ./microbench
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define LOADV uint32_t v = s[(it + threadIdx.x / 32) & 255];

// ---- variant 0: ISETP.EQ.OR, 32 rows ------------------------------------------------------------
__global__ void __launch_bounds__(256) k_isetp(const uint32_t* in, uint32_t* out, int iters) {
    __shared__ uint32_t s[256];
    s[threadIdx.x] = in[1024 + threadIdx.x];
    uint32_t r[32];
    #pragma unroll
    for (int q = 0; q < 32; ++q) r[q] = in[(threadIdx.x + 33 * q) & 1023];
    __syncthreads();
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
        LOADV
        uint32_t f;
        asm volatile("{\n.reg .pred p0,p1,p2,p3;\n"
            "setp.eq.u32 p0, %1, %33;\n setp.eq.u32 p1, %2, %33;\n setp.eq.u32 p2, %3, %33;\n setp.eq.u32 p3, %4, %33;\n"
            "setp.eq.or.u32 p0, %5, %33, p0;\n setp.eq.or.u32 p1, %6, %33, p1;\n setp.eq.or.u32 p2, %7, %33, p2;\n setp.eq.or.u32 p3, %8, %33, p3;\n"
            "setp.eq.or.u32 p0, %9, %33, p0;\n setp.eq.or.u32 p1, %10, %33, p1;\n setp.eq.or.u32 p2, %11, %33, p2;\n setp.eq.or.u32 p3, %12, %33, p3;\n"
            "setp.eq.or.u32 p0, %13, %33, p0;\n setp.eq.or.u32 p1, %14, %33, p1;\n setp.eq.or.u32 p2, %15, %33, p2;\n setp.eq.or.u32 p3, %16, %33, p3;\n"
            "setp.eq.or.u32 p0, %17, %33, p0;\n setp.eq.or.u32 p1, %18, %33, p1;\n setp.eq.or.u32 p2, %19, %33, p2;\n setp.eq.or.u32 p3, %20, %33, p3;\n"
            "setp.eq.or.u32 p0, %21, %33, p0;\n setp.eq.or.u32 p1, %22, %33, p1;\n setp.eq.or.u32 p2, %23, %33, p2;\n setp.eq.or.u32 p3, %24, %33, p3;\n"
            "setp.eq.or.u32 p0, %25, %33, p0;\n setp.eq.or.u32 p1, %26, %33, p1;\n setp.eq.or.u32 p2, %27, %33, p2;\n setp.eq.or.u32 p3, %28, %33, p3;\n"
            "setp.eq.or.u32 p0, %29, %33, p0;\n setp.eq.or.u32 p1, %30, %33, p1;\n setp.eq.or.u32 p2, %31, %33, p2;\n setp.eq.or.u32 p3, %32, %33, p3;\n"
            "or.pred p0, p0, p1;\n or.pred p2, p2, p3;\n or.pred p0, p0, p2;\n selp.u32 %0, 1, 0, p0;\n}\n"
            : "=r"(f) : "r"(r[0]),"r"(r[1]),"r"(r[2]),"r"(r[3]),"r"(r[4]),"r"(r[5]),"r"(r[6]),"r"(r[7]),"r"(r[8]),"r"(r[9]),"r"(r[10]),"r"(r[11]),
              "r"(r[12]),"r"(r[13]),"r"(r[14]),"r"(r[15]),"r"(r[16]),"r"(r[17]),"r"(r[18]),"r"(r[19]),"r"(r[20]),"r"(r[21]),"r"(r[22]),"r"(r[23]),
              "r"(r[24]),"r"(r[25]),"r"(r[26]),"r"(r[27]),"r"(r[28]),"r"(r[29]),"r"(r[30]),"r"(r[31]),"r"(v));
        acc |= f;
    }
    if (acc) out[0] = acc;
}

// ---- variant 1: FSETP.EQ.OR, 32 rows (bit patterns as floats) ------------------------------------
__global__ void __launch_bounds__(256) k_fsetp(const uint32_t* in, uint32_t* out, int iters) {
    __shared__ uint32_t s[256];
    s[threadIdx.x] = in[1024 + threadIdx.x];
    uint32_t r[32];
    #pragma unroll
    for (int q = 0; q < 32; ++q) r[q] = in[(threadIdx.x + 33 * q) & 1023];
    __syncthreads();
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
        LOADV
        uint32_t f;
        asm volatile("{\n.reg .pred p0,p1,p2,p3;\n.reg .f32 fv;\n mov.b32 fv, %33;\n"
            "setp.eq.f32 p0, %1, fv;\n setp.eq.f32 p1, %2, fv;\n setp.eq.f32 p2, %3, fv;\n setp.eq.f32 p3, %4, fv;\n"
            "setp.eq.or.f32 p0, %5, fv, p0;\n setp.eq.or.f32 p1, %6, fv, p1;\n setp.eq.or.f32 p2, %7, fv, p2;\n setp.eq.or.f32 p3, %8, fv, p3;\n"
            "setp.eq.or.f32 p0, %9, fv, p0;\n setp.eq.or.f32 p1, %10, fv, p1;\n setp.eq.or.f32 p2, %11, fv, p2;\n setp.eq.or.f32 p3, %12, fv, p3;\n"
            "setp.eq.or.f32 p0, %13, fv, p0;\n setp.eq.or.f32 p1, %14, fv, p1;\n setp.eq.or.f32 p2, %15, fv, p2;\n setp.eq.or.f32 p3, %16, fv, p3;\n"
            "setp.eq.or.f32 p0, %17, fv, p0;\n setp.eq.or.f32 p1, %18, fv, p1;\n setp.eq.or.f32 p2, %19, fv, p2;\n setp.eq.or.f32 p3, %20, fv, p3;\n"
            "setp.eq.or.f32 p0, %21, fv, p0;\n setp.eq.or.f32 p1, %22, fv, p1;\n setp.eq.or.f32 p2, %23, fv, p2;\n setp.eq.or.f32 p3, %24, fv, p3;\n"
            "setp.eq.or.f32 p0, %25, fv, p0;\n setp.eq.or.f32 p1, %26, fv, p1;\n setp.eq.or.f32 p2, %27, fv, p2;\n setp.eq.or.f32 p3, %28, fv, p3;\n"
            "setp.eq.or.f32 p0, %29, fv, p0;\n setp.eq.or.f32 p1, %30, fv, p1;\n setp.eq.or.f32 p2, %31, fv, p2;\n setp.eq.or.f32 p3, %32, fv, p3;\n"
            "or.pred p0, p0, p1;\n or.pred p2, p2, p3;\n or.pred p0, p0, p2;\n selp.u32 %0, 1, 0, p0;\n}\n"
            : "=r"(f) : "f"(__uint_as_float(r[0])),"f"(__uint_as_float(r[1])),"f"(__uint_as_float(r[2])),"f"(__uint_as_float(r[3])),
              "f"(__uint_as_float(r[4])),"f"(__uint_as_float(r[5])),"f"(__uint_as_float(r[6])),"f"(__uint_as_float(r[7])),
              "f"(__uint_as_float(r[8])),"f"(__uint_as_float(r[9])),"f"(__uint_as_float(r[10])),"f"(__uint_as_float(r[11])),
              "f"(__uint_as_float(r[12])),"f"(__uint_as_float(r[13])),"f"(__uint_as_float(r[14])),"f"(__uint_as_float(r[15])),
              "f"(__uint_as_float(r[16])),"f"(__uint_as_float(r[17])),"f"(__uint_as_float(r[18])),"f"(__uint_as_float(r[19])),
              "f"(__uint_as_float(r[20])),"f"(__uint_as_float(r[21])),"f"(__uint_as_float(r[22])),"f"(__uint_as_float(r[23])),
              "f"(__uint_as_float(r[24])),"f"(__uint_as_float(r[25])),"f"(__uint_as_float(r[26])),"f"(__uint_as_float(r[27])),
              "f"(__uint_as_float(r[28])),"f"(__uint_as_float(r[29])),"f"(__uint_as_float(r[30])),"f"(__uint_as_float(r[31])),"r"(v));
        acc |= f;
    }
    if (acc) out[0] = acc;
}

// ---- variant 2: HSET2 (bitmask) + LOP3: 16 half2 registers = 32 rows, 2 rows per compare ---------
// set.ne.u32.f16x2 gives 0xFFFF per half that differs; AND-accumulate: a zero half at the end = a match.
#define HSET2_BLOCK(NM) \
        asm volatile("{\n.reg .b32 t0,t1,t2,t3;\n" \
            "set.ne.u32.f16x2 t0, %1, %17;\n set.ne.u32.f16x2 t1, %2, %17;\n and.b32 t0, t0, t1;\n" \
            "set.ne.u32.f16x2 t2, %3, %17;\n set.ne.u32.f16x2 t3, %4, %17;\n and.b32 t2, t2, t3;\n" \
            "set.ne.u32.f16x2 t1, %5, %17;\n and.b32 t0, t0, t1;\n set.ne.u32.f16x2 t3, %6, %17;\n and.b32 t2, t2, t3;\n" \
            "set.ne.u32.f16x2 t1, %7, %17;\n and.b32 t0, t0, t1;\n set.ne.u32.f16x2 t3, %8, %17;\n and.b32 t2, t2, t3;\n" \
            "set.ne.u32.f16x2 t1, %9, %17;\n and.b32 t0, t0, t1;\n set.ne.u32.f16x2 t3, %10, %17;\n and.b32 t2, t2, t3;\n" \
            "set.ne.u32.f16x2 t1, %11, %17;\n and.b32 t0, t0, t1;\n set.ne.u32.f16x2 t3, %12, %17;\n and.b32 t2, t2, t3;\n" \
            "set.ne.u32.f16x2 t1, %13, %17;\n and.b32 t0, t0, t1;\n set.ne.u32.f16x2 t3, %14, %17;\n and.b32 t2, t2, t3;\n" \
            "set.ne.u32.f16x2 t1, %15, %17;\n and.b32 t0, t0, t1;\n set.ne.u32.f16x2 t3, %16, %17;\n and.b32 t2, t2, t3;\n" \
            "and.b32 %0, t0, t2;\n}\n" \
            : "=r"(NM) : "r"(h[0]),"r"(h[1]),"r"(h[2]),"r"(h[3]),"r"(h[4]),"r"(h[5]),"r"(h[6]),"r"(h[7]),"r"(h[8]),"r"(h[9]),"r"(h[10]),"r"(h[11]), \
              "r"(h[12]),"r"(h[13]),"r"(h[14]),"r"(h[15]),"r"(v));

__global__ void __launch_bounds__(256) k_hset2(const uint32_t* in, uint32_t* out, int iters) {
    __shared__ uint32_t s[256];
    s[threadIdx.x] = in[1024 + threadIdx.x];
    uint32_t h[16];
    #pragma unroll
    for (int q = 0; q < 16; ++q) h[q] = in[(threadIdx.x + 33 * q) & 1023] & 0x7BFF7BFFu;
    __syncthreads();
    uint32_t acc = 0xFFFFFFFFu;
    for (int it = 0; it < iters; ++it) {
        LOADV
        v &= 0x7BFF7BFFu;
        uint32_t nm;
        HSET2_BLOCK(nm)
        acc &= nm;
    }
    if (acc != 0xFFFFFFFFu) out[0] = acc;
}

// ---- variant 3: HSETP2 + PLOP3 (predicate accumulate) --------------------------------------------
__global__ void __launch_bounds__(256) k_hsetp2(const uint32_t* in, uint32_t* out, int iters) {
    __shared__ uint32_t s[256];
    s[threadIdx.x] = in[1024 + threadIdx.x];
    uint32_t h[16];
    #pragma unroll
    for (int q = 0; q < 16; ++q) h[q] = in[(threadIdx.x + 33 * q) & 1023] & 0x7BFF7BFFu;
    __syncthreads();
    uint32_t acc = 1;
    for (int it = 0; it < iters; ++it) {
        LOADV
        v &= 0x7BFF7BFFu;
        uint32_t f;
        asm volatile("{\n.reg .pred a, b, p, q;\n setp.ne.u32 a, %17, 0xFFFFFFFF;\n setp.ne.u32 b, %17, 0xFFFFFFFF;\n"
            "setp.ne.f16x2 p|q, %1, %17;\n and.pred a, a, p;\n and.pred a, a, q;\n setp.ne.f16x2 p|q, %2, %17;\n and.pred b, b, p;\n and.pred b, b, q;\n"
            "setp.ne.f16x2 p|q, %3, %17;\n and.pred a, a, p;\n and.pred a, a, q;\n setp.ne.f16x2 p|q, %4, %17;\n and.pred b, b, p;\n and.pred b, b, q;\n"
            "setp.ne.f16x2 p|q, %5, %17;\n and.pred a, a, p;\n and.pred a, a, q;\n setp.ne.f16x2 p|q, %6, %17;\n and.pred b, b, p;\n and.pred b, b, q;\n"
            "setp.ne.f16x2 p|q, %7, %17;\n and.pred a, a, p;\n and.pred a, a, q;\n setp.ne.f16x2 p|q, %8, %17;\n and.pred b, b, p;\n and.pred b, b, q;\n"
            "setp.ne.f16x2 p|q, %9, %17;\n and.pred a, a, p;\n and.pred a, a, q;\n setp.ne.f16x2 p|q, %10, %17;\n and.pred b, b, p;\n and.pred b, b, q;\n"
            "setp.ne.f16x2 p|q, %11, %17;\n and.pred a, a, p;\n and.pred a, a, q;\n setp.ne.f16x2 p|q, %12, %17;\n and.pred b, b, p;\n and.pred b, b, q;\n"
            "setp.ne.f16x2 p|q, %13, %17;\n and.pred a, a, p;\n and.pred a, a, q;\n setp.ne.f16x2 p|q, %14, %17;\n and.pred b, b, p;\n and.pred b, b, q;\n"
            "setp.ne.f16x2 p|q, %15, %17;\n and.pred a, a, p;\n and.pred a, a, q;\n setp.ne.f16x2 p|q, %16, %17;\n and.pred b, b, p;\n and.pred b, b, q;\n"
            "and.pred a, a, b;\n selp.u32 %0, 1, 0, a;\n}\n"
            : "=r"(f) : "r"(h[0]),"r"(h[1]),"r"(h[2]),"r"(h[3]),"r"(h[4]),"r"(h[5]),"r"(h[6]),"r"(h[7]),"r"(h[8]),"r"(h[9]),"r"(h[10]),"r"(h[11]),
              "r"(h[12]),"r"(h[13]),"r"(h[14]),"r"(h[15]),"r"(v));
        acc &= f;
    }
    if (!acc) out[0] = 7;
}

// ---- variant 4: mix: 16 half2 registers through HSET2+LOP3 and 8 u32 registers through ISETP -------
__global__ void __launch_bounds__(256) k_mix_hset2_isetp(const uint32_t* in, uint32_t* out, int iters) {
    __shared__ uint32_t s[256];
    __shared__ uint32_t s2[256];
    s[threadIdx.x] = in[1024 + threadIdx.x];
    s2[threadIdx.x] = in[1280 + threadIdx.x];
    uint32_t h[16], r[8];
    #pragma unroll
    for (int q = 0; q < 16; ++q) h[q] = in[(threadIdx.x + 33 * q) & 1023] & 0x7BFF7BFFu;
    #pragma unroll
    for (int q = 0; q < 8; ++q) r[q] = in[(threadIdx.x + 37 * q + 5) & 1023];
    __syncthreads();
    uint32_t acc = 0xFFFFFFFFu, acc2 = 0;
    for (int it = 0; it < iters; ++it) {
        LOADV
        uint32_t w = s2[(it + threadIdx.x / 32) & 255];
        v &= 0x7BFF7BFFu;
        uint32_t nm, f;
        HSET2_BLOCK(nm)
        asm volatile("{\n.reg .pred p0,p1;\n"
            "setp.eq.u32 p0, %1, %9;\n setp.eq.u32 p1, %2, %9;\n setp.eq.or.u32 p0, %3, %9, p0;\n setp.eq.or.u32 p1, %4, %9, p1;\n"
            "setp.eq.or.u32 p0, %5, %9, p0;\n setp.eq.or.u32 p1, %6, %9, p1;\n setp.eq.or.u32 p0, %7, %9, p0;\n setp.eq.or.u32 p1, %8, %9, p1;\n"
            "or.pred p0, p0, p1;\n selp.u32 %0, 1, 0, p0;\n}\n"
            : "=r"(f) : "r"(r[0]),"r"(r[1]),"r"(r[2]),"r"(r[3]),"r"(r[4]),"r"(r[5]),"r"(r[6]),"r"(r[7]),"r"(w));
        acc &= nm; acc2 |= f;
    }
    if (acc != 0xFFFFFFFFu || acc2) out[0] = acc + acc2;
}

// ---- variant 5: IMAD chain (FMA pipe integer rate) ------------------------------------------------
__global__ void __launch_bounds__(256) k_imad(const uint32_t* in, uint32_t* out, int iters) {
    __shared__ uint32_t s[256];
    s[threadIdx.x] = in[1024 + threadIdx.x];
    uint32_t c[32];
    #pragma unroll
    for (int q = 0; q < 32; ++q) c[q] = in[(threadIdx.x + 33 * q) & 1023];
    __syncthreads();
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
        LOADV
        uint32_t a0 = c[0], a1 = c[8], a2 = c[16], a3 = c[24];
        #pragma unroll
        for (int q = 1; q < 8; ++q) { a0 = a0 * v + c[q]; a1 = a1 * v + c[8 + q]; a2 = a2 * v + c[16 + q]; a3 = a3 * v + c[24 + q]; }
        acc ^= a0 ^ a1 ^ a2 ^ a3;
    }
    if (acc == 0x1234567) out[0] = acc;
}

// ---- variant 6: LOP3 only -------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_lop3(const uint32_t* in, uint32_t* out, int iters) {
    __shared__ uint32_t s[256];
    s[threadIdx.x] = in[1024 + threadIdx.x];
    uint32_t r[32];
    #pragma unroll
    for (int q = 0; q < 32; ++q) r[q] = in[(threadIdx.x + 33 * q) & 1023];
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
        LOADV
        #pragma unroll
        for (int q = 0; q < 32; ++q) r[q] = (r[q] ^ v) & (r[(q + 1) & 31] | v);
    }
    uint32_t a = 0;
    #pragma unroll
    for (int q = 0; q < 32; ++q) a ^= r[q];
    if (a == 0x12345678u) out[0] = a;
}


// ---- variant 7: proposed dual-pipe inner loop: 16 rows by ISETP (ALU pipe) + 16 rows as two monic degree-8
// polynomials P(v) = prod (v - r_q) mod 2^32 evaluated by Horner with IMAD (FMA pipe), 2 zero tests --------
__global__ void __launch_bounds__(128) k_mix_isetp_horner(const uint32_t* in, uint32_t* out, int iters) {
    __shared__ uint32_t s[256];
    s[threadIdx.x] = in[1024 + threadIdx.x]; s[threadIdx.x + 128] = in[1152 + threadIdx.x];
    uint32_t r[16], c[16];
    #pragma unroll
    for (int q = 0; q < 16; ++q) { r[q] = in[(threadIdx.x + 33 * q) & 1023]; c[q] = in[(threadIdx.x + 35 * q + 7) & 1023]; }
    __syncthreads();
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
        LOADV
        uint32_t f;
        asm volatile("{\n.reg .pred p0,p1,p2,p3;\n.reg .u32 a, b;\n"
            "mad.lo.u32 a, %17, 1, %18;\n mad.lo.u32 b, %17, 1, %26;\n"
            "setp.eq.u32 p0, %1, %17;\n setp.eq.u32 p1, %2, %17;\n"
            "mad.lo.u32 a, a, %17, %19;\n mad.lo.u32 b, b, %17, %27;\n"
            "setp.eq.u32 p2, %3, %17;\n setp.eq.u32 p3, %4, %17;\n"
            "mad.lo.u32 a, a, %17, %20;\n mad.lo.u32 b, b, %17, %28;\n"
            "setp.eq.or.u32 p0, %5, %17, p0;\n setp.eq.or.u32 p1, %6, %17, p1;\n"
            "mad.lo.u32 a, a, %17, %21;\n mad.lo.u32 b, b, %17, %29;\n"
            "setp.eq.or.u32 p2, %7, %17, p2;\n setp.eq.or.u32 p3, %8, %17, p3;\n"
            "mad.lo.u32 a, a, %17, %22;\n mad.lo.u32 b, b, %17, %30;\n"
            "setp.eq.or.u32 p0, %9, %17, p0;\n setp.eq.or.u32 p1, %10, %17, p1;\n"
            "mad.lo.u32 a, a, %17, %23;\n mad.lo.u32 b, b, %17, %31;\n"
            "setp.eq.or.u32 p2, %11, %17, p2;\n setp.eq.or.u32 p3, %12, %17, p3;\n"
            "mad.lo.u32 a, a, %17, %24;\n mad.lo.u32 b, b, %17, %32;\n"
            "setp.eq.or.u32 p0, %13, %17, p0;\n setp.eq.or.u32 p1, %14, %17, p1;\n"
            "mad.lo.u32 a, a, %17, %25;\n mad.lo.u32 b, b, %17, %33;\n"
            "setp.eq.or.u32 p2, %15, %17, p2;\n setp.eq.or.u32 p3, %16, %17, p3;\n"
            "setp.eq.or.u32 p0, a, 0, p0;\n setp.eq.or.u32 p1, b, 0, p1;\n"
            "or.pred p0, p0, p1;\n or.pred p2, p2, p3;\n or.pred p0, p0, p2;\n selp.u32 %0, 1, 0, p0;\n}\n"
            : "=r"(f) : "r"(r[0]),"r"(r[1]),"r"(r[2]),"r"(r[3]),"r"(r[4]),"r"(r[5]),"r"(r[6]),"r"(r[7]),"r"(r[8]),"r"(r[9]),"r"(r[10]),"r"(r[11]),
              "r"(r[12]),"r"(r[13]),"r"(r[14]),"r"(r[15]),"r"(v),
              "r"(c[0]),"r"(c[1]),"r"(c[2]),"r"(c[3]),"r"(c[4]),"r"(c[5]),"r"(c[6]),"r"(c[7]),"r"(c[8]),"r"(c[9]),"r"(c[10]),"r"(c[11]),
              "r"(c[12]),"r"(c[13]),"r"(c[14]),"r"(c[15]));
        acc |= f;
    }
    if (acc) out[0] = acc;
}

// ---- variant 8: integer issue peak: independent LOP3 (ALU pipe) and IMAD (FMA pipe) streams ------------
__global__ void __launch_bounds__(256) k_mix_lop3_imad(const uint32_t* in, uint32_t* out, int iters) {
    __shared__ uint32_t s[256];
    s[threadIdx.x] = in[1024 + threadIdx.x];
    uint32_t r[16], c[16];
    #pragma unroll
    for (int q = 0; q < 16; ++q) { r[q] = in[(threadIdx.x + 33 * q) & 1023]; c[q] = in[(threadIdx.x + 35 * q + 7) & 1023]; }
    __syncthreads();
    for (int it = 0; it < iters; ++it) {
        LOADV
        #pragma unroll
        for (int q = 0; q < 16; ++q) { r[q] = (r[q] ^ v) & (r[(q + 1) & 15] | v); c[q] = c[q] * v + c[(q + 3) & 15]; }
    }
    uint32_t a = 0;
    #pragma unroll
    for (int q = 0; q < 16; ++q) a ^= r[q] ^ c[q];
    if (a == 0x12345678u) out[0] = a;
}

template <typename K>
double run(const char* name, K kern, const uint32_t* in, uint32_t* out, double cells_per_iter, int sm) {
    const int grid = sm * 8, iters = 1 << 14;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a);
        kern<<<grid, 256>>>(in, out, iters);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0; cudaEventElapsedTime(&ms, a, b);
        double rate = (double)grid * 256.0 * iters * cells_per_iter / (ms * 1e-3);
        if (rep) best = rate > best ? rate : best;
    }
    cudaError_t e = cudaGetLastError();
    printf("%-28s %.4e cells/s  (%.1f per clk per SM at 1.9 GHz)  %s\n", name, best, best / sm / 1.9e9, e == cudaSuccess ? "" : cudaGetErrorString(e));
    return best;
}

template <typename K>
double run128(const char* name, K kern, const uint32_t* in, uint32_t* out, double cells_per_iter, int sm) {
    const int grid = sm * 16, iters = 1 << 14;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a);
        kern<<<grid, 128>>>(in, out, iters);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0; cudaEventElapsedTime(&ms, a, b);
        double rate = (double)grid * 128.0 * iters * cells_per_iter / (ms * 1e-3);
        if (rep) best = rate > best ? rate : best;
    }
    cudaError_t e = cudaGetLastError();
    printf("%-28s %.4e cells/s  (%.1f per clk per SM at 1.9 GHz)  %s\n", name, best, best / sm / 1.9e9, e == cudaSuccess ? "" : cudaGetErrorString(e));
    return best;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sm = p.multiProcessorCount;
    printf("%s, %d SMs, clock %d kHz\n", p.name, sm, p.clockRate);
    uint32_t h[2048];
    uint32_t x = 12345;
    for (int i = 0; i < 2048; ++i) { x = x * 1664525u + 1013904223u; h[i] = x | 1u; }
    uint32_t *in, *out; cudaMalloc(&in, sizeof(h)); cudaMalloc(&out, 64); cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    run("isetp32 (ISETP.EQ.OR)", k_isetp, in, out, 32, sm);
    run("fsetp32 (FSETP.EQ.OR)", k_fsetp, in, out, 32, sm);
    run("hset2+lop3 (32 rows)", k_hset2, in, out, 32, sm);
    run("hsetp2+plop3 (32 rows)", k_hsetp2, in, out, 32, sm);
    run("mix hset2(32)+isetp(8)", k_mix_hset2_isetp, in, out, 40, sm);
    run("imad (28 per iter)", k_imad, in, out, 28, sm);
    run("lop3 (32 per iter)", k_lop3, in, out, 32, sm);
    run("mix lop3(16)+imad(16)", k_mix_lop3_imad, in, out, 32, sm);
    run128("mix isetp(16)+horner(16)", k_mix_isetp_horner, in, out, 32, sm);
    return 0;
}
