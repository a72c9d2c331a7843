// Inner-loop candidates for the tile kernel, in the real loop shape (sm_100a):
// a warp holds its read rows in registers and walks a shared-memory buffer of structure words in
// vote blocks of 32; one ballot per block.  Prints cells/s per variant.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench2 tools/microbench2.cu && ./microbench2
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int TS = 2048;          // structure words per buffer

__device__ __forceinline__ uint32_t hne2(uint32_t a, uint32_t b) {      // 0xFFFF per half that differs (f16x2 compare)
    uint32_t d;
    asm("set.ne.u32.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t and3(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0x80;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t dup_lo(uint32_t w) { return __byte_perm(w, 0, 0x1010); }
__device__ __forceinline__ uint32_t dup_hi(uint32_t w) { return __byte_perm(w, 0, 0x3232); }

// ---- V0: R ISETP rows against full words -----------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(128) k_isetp(const uint32_t* in, uint32_t* out, int reps) {
    __shared__ __align__(16) uint32_t s[4][TS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = lane; i < TS; i += 32) s[warp][i] = in[(i * 7 + warp) & 4095];
    uint32_t r[R];
    #pragma unroll
    for (int q = 0; q < R; ++q) r[q] = in[(threadIdx.x * R + q + blockIdx.x) & 4095] | 0x40000000u;
    __syncwarp();
    uint32_t found = 0;
    for (int rep = 0; rep < reps; ++rep) {
        for (int b = 0; b < TS / 32; ++b) {
            const uint32_t* sblk = s[warp] + b * 32;
            bool p0 = false, p1 = false, p2 = false, p3 = false;
            #pragma unroll
            for (int jj = 0; jj < 32; jj += 4) {
                const uint4 v = *reinterpret_cast<const uint4*>(sblk + jj);
                #pragma unroll
                for (int q = 0; q < R; ++q) {
                    p0 |= (r[q] == v.x); p1 |= (r[q] == v.y); p2 |= (r[q] == v.z); p3 |= (r[q] == v.w);
                }
            }
            unsigned mask = __ballot_sync(0xFFFFFFFFu, p0 | p1 | p2 | p3);
            if (mask) found += __popc(mask);
        }
    }
    if (found) out[0] = found;
}

// ---- V1: NH half2 registers (2*NH rows) through HSET2 + 3-input LOP3, 16-bit filter words in shared ---
template <int NH>
__global__ void __launch_bounds__(128) k_hset2(const uint32_t* in, uint32_t* out, int reps) {
    __shared__ __align__(16) uint16_t s[4][TS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = lane; i < TS; i += 32) s[warp][i] = (uint16_t)(in[(i * 7 + warp) & 4095] & 0x3BFF);
    uint32_t h[NH];
    #pragma unroll
    for (int q = 0; q < NH; ++q) h[q] = (in[(threadIdx.x * NH + q + blockIdx.x) & 4095] & 0x3BFF3BFFu) | 0x40004000u;
    __syncwarp();
    uint32_t found = 0;
    for (int rep = 0; rep < reps; ++rep) {
        for (int b = 0; b < TS / 32; ++b) {
            const uint16_t* sblk = s[warp] + b * 32;
            uint32_t a0 = 0xFFFFFFFFu, a1 = 0xFFFFFFFFu, a2 = 0xFFFFFFFFu, a3 = 0xFFFFFFFFu;
            #pragma unroll
            for (int jj = 0; jj < 32; jj += 8) {
                const uint4 v = *reinterpret_cast<const uint4*>(sblk + jj);
                const uint32_t vv[8] = {dup_lo(v.x), dup_hi(v.x), dup_lo(v.y), dup_hi(v.y), dup_lo(v.z), dup_hi(v.z), dup_lo(v.w), dup_hi(v.w)};
                #pragma unroll
                for (int e = 0; e < 8; ++e) {
                    #pragma unroll
                    for (int q = 0; q < NH; q += 8) {
                        a0 = and3(a0, hne2(h[q], vv[e]), hne2(h[q + 1], vv[e]));
                        a1 = and3(a1, hne2(h[q + 2], vv[e]), hne2(h[q + 3], vv[e]));
                        a2 = and3(a2, hne2(h[q + 4], vv[e]), hne2(h[q + 5], vv[e]));
                        a3 = and3(a3, hne2(h[q + 6], vv[e]), hne2(h[q + 7], vv[e]));
                    }
                }
            }
            unsigned mask = __ballot_sync(0xFFFFFFFFu, and3(a0, a1, a2 & a3) != 0xFFFFFFFFu);
            if (mask) found += __popc(mask);
        }
    }
    if (found) out[0] = found;
}

// ---- V2: mixed: NH half2 registers through HSET2 (fma pipe) + NI full-word ISETP rows (alu pipe) -----
// one shared array of full words; the 16-bit filter is the low half of the word (dup by PRMT)
template <int NH, int NI>
__global__ void __launch_bounds__(128) k_mix(const uint32_t* in, uint32_t* out, int reps) {
    __shared__ __align__(16) uint32_t s[4][TS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = lane; i < TS; i += 32) s[warp][i] = in[(i * 7 + warp) & 4095] & 0x3FFF3BFFu;
    uint32_t h[NH], r[NI];
    #pragma unroll
    for (int q = 0; q < NH; ++q) h[q] = (in[(threadIdx.x * NH + q + blockIdx.x) & 4095] & 0x3BFF3BFFu) | 0x40004000u;
    #pragma unroll
    for (int q = 0; q < NI; ++q) r[q] = in[(threadIdx.x * NI + q + 3 * blockIdx.x) & 4095] | 0x40000000u;
    __syncwarp();
    uint32_t found = 0;
    for (int rep = 0; rep < reps; ++rep) {
        for (int b = 0; b < TS / 32; ++b) {
            const uint32_t* sblk = s[warp] + b * 32;
            uint32_t a0 = 0xFFFFFFFFu, a1 = 0xFFFFFFFFu, a2 = 0xFFFFFFFFu, a3 = 0xFFFFFFFFu;
            bool p0 = false, p1 = false;
            #pragma unroll
            for (int jj = 0; jj < 32; jj += 4) {
                const uint4 v = *reinterpret_cast<const uint4*>(sblk + jj);
                const uint32_t vw[4] = {v.x, v.y, v.z, v.w};
                #pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint32_t vd = dup_lo(vw[e]);
                    #pragma unroll
                    for (int q = 0; q < NH; q += 8) {
                        a0 = and3(a0, hne2(h[q], vd), hne2(h[q + 1], vd));
                        a1 = and3(a1, hne2(h[q + 2], vd), hne2(h[q + 3], vd));
                        a2 = and3(a2, hne2(h[q + 4], vd), hne2(h[q + 5], vd));
                        a3 = and3(a3, hne2(h[q + 6], vd), hne2(h[q + 7], vd));
                    }
                    #pragma unroll
                    for (int q = 0; q < NI; q += 2) { p0 |= (r[q] == vw[e]); p1 |= (r[q + 1] == vw[e]); }
                }
            }
            unsigned mask = __ballot_sync(0xFFFFFFFFu, (and3(a0, a1, a2 & a3) != 0xFFFFFFFFu) | p0 | p1);
            if (mask) found += __popc(mask);
        }
    }
    if (found) out[0] = found;
}

// ---- V3: mixed ISETP rows + monic degree-D polynomials by Horner (IMAD on the fma pipe) ---------------
template <int NI, int NP, int D>
__global__ void __launch_bounds__(128) k_horner(const uint32_t* in, uint32_t* out, int reps) {
    __shared__ __align__(16) uint32_t s[4][TS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = lane; i < TS; i += 32) s[warp][i] = in[(i * 7 + warp) & 4095] | 1u;
    uint32_t r[NI], c[NP][D];
    #pragma unroll
    for (int q = 0; q < NI; ++q) r[q] = in[(threadIdx.x * NI + q + 3 * blockIdx.x) & 4095] & ~1u;
    #pragma unroll
    for (int p = 0; p < NP; ++p)
        #pragma unroll
        for (int q = 0; q < D; ++q) c[p][q] = in[(threadIdx.x * 16 + p * D + q + 5 * blockIdx.x) & 4095] | 1u;
    __syncwarp();
    uint32_t found = 0;
    for (int rep = 0; rep < reps; ++rep) {
        for (int b = 0; b < TS / 32; ++b) {
            const uint32_t* sblk = s[warp] + b * 32;
            bool p0 = false, p1 = false, p2 = false, p3 = false;
            #pragma unroll
            for (int jj = 0; jj < 32; jj += 4) {
                const uint4 v = *reinterpret_cast<const uint4*>(sblk + jj);
                const uint32_t vw[4] = {v.x, v.y, v.z, v.w};
                #pragma unroll
                for (int e = 0; e < 4; ++e) {
                    uint32_t acc[NP];
                    #pragma unroll
                    for (int p = 0; p < NP; ++p) acc[p] = vw[e] + c[p][0];
                    #pragma unroll
                    for (int q = 1; q < D; ++q)
                        #pragma unroll
                        for (int p = 0; p < NP; ++p) acc[p] = acc[p] * vw[e] + c[p][q];
                    #pragma unroll
                    for (int q = 0; q < NI; q += 2) { p0 |= (r[q] == vw[e]); p1 |= (r[q + 1] == vw[e]); }
                    #pragma unroll
                    for (int p = 0; p < NP; ++p) { if (p & 1) p3 |= (acc[p] == 0); else p2 |= (acc[p] == 0); }
                }
            }
            unsigned mask = __ballot_sync(0xFFFFFFFFu, p0 | p1 | p2 | p3);
            if (mask) found += __popc(mask);
        }
    }
    if (found) out[0] = found;
}

// ---- V4 (stream word first in every instruction: operand-reuse cache friendly): mixed ISETP rows + monic degree-D polynomials by Horner (IMAD on the fma pipe) ---------------
template <int NI, int NP, int D>
__global__ void __launch_bounds__(128) k_horner_a(const uint32_t* in, uint32_t* out, int reps) {
    __shared__ __align__(16) uint32_t s[4][TS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = lane; i < TS; i += 32) s[warp][i] = in[(i * 7 + warp) & 4095] | 1u;
    uint32_t r[NI], c[NP][D];
    #pragma unroll
    for (int q = 0; q < NI; ++q) r[q] = in[(threadIdx.x * NI + q + 3 * blockIdx.x) & 4095] & ~1u;
    #pragma unroll
    for (int p = 0; p < NP; ++p)
        #pragma unroll
        for (int q = 0; q < D; ++q) c[p][q] = in[(threadIdx.x * 16 + p * D + q + 5 * blockIdx.x) & 4095] | 1u;
    __syncwarp();
    uint32_t found = 0;
    for (int rep = 0; rep < reps; ++rep) {
        for (int b = 0; b < TS / 32; ++b) {
            const uint32_t* sblk = s[warp] + b * 32;
            bool p0 = false, p1 = false, p2 = false, p3 = false;
            #pragma unroll
            for (int jj = 0; jj < 32; jj += 4) {
                const uint4 v = *reinterpret_cast<const uint4*>(sblk + jj);
                const uint32_t vw[4] = {v.x, v.y, v.z, v.w};
                #pragma unroll
                for (int e = 0; e < 4; ++e) {
                    uint32_t acc[NP];
                    #pragma unroll
                    for (int p = 0; p < NP; ++p) acc[p] = vw[e] + c[p][0];
                    #pragma unroll
                    for (int q = 1; q < D; ++q)
                        #pragma unroll
                        for (int p = 0; p < NP; ++p) acc[p] = vw[e] * acc[p] + c[p][q];
                    #pragma unroll
                    for (int q = 0; q < NI; q += 2) { p0 |= (vw[e] == r[q]); p1 |= (vw[e] == r[q + 1]); }
                    #pragma unroll
                    for (int p = 0; p < NP; ++p) { if (p & 1) p3 |= (acc[p] == 0); else p2 |= (acc[p] == 0); }
                }
            }
            unsigned mask = __ballot_sync(0xFFFFFFFFu, p0 | p1 | p2 | p3);
            if (mask) found += __popc(mask);
        }
    }
    if (found) out[0] = found;
}

// ---- V5: NH packed half2 registers (2*NH rows, 16-bit filters) through HSET2 + 3-input LOP3 on the alu pipe
//           + NP degree-D row polynomials by Horner on the fma pipe; streamed word first in every instruction ----
template <int NH, int NP, int D>
__global__ void __launch_bounds__(128) k_hset2_horner(const uint32_t* in, uint32_t* out, int reps) {
    __shared__ __align__(16) uint32_t s[4][TS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = lane; i < TS; i += 32) s[warp][i] = (in[(i * 7 + warp) & 4095] & 0xFFFFBBFFu) | 1u;
    uint32_t h[NH], c[NP][D];
    #pragma unroll
    for (int q = 0; q < NH; ++q) h[q] = (in[(threadIdx.x * NH + q + blockIdx.x) & 4095] & 0x3BFE3BFEu) | 0x40004000u;
    #pragma unroll
    for (int p = 0; p < NP; ++p)
        #pragma unroll
        for (int q = 0; q < D; ++q) c[p][q] = in[(threadIdx.x * 16 + p * D + q + 5 * blockIdx.x) & 4095] | 1u;
    __syncwarp();
    uint32_t found = 0;
    for (int rep = 0; rep < reps; ++rep) {
        for (int b = 0; b < TS / 32; ++b) {
            const uint32_t* sblk = s[warp] + b * 32;
            uint32_t a0 = 0xFFFFFFFFu, a1 = 0xFFFFFFFFu;
            bool p2 = false, p3 = false;
            #pragma unroll
            for (int jj = 0; jj < 32; jj += 4) {
                const uint4 v4 = *reinterpret_cast<const uint4*>(sblk + jj);
                const uint32_t vw[4] = {v4.x, v4.y, v4.z, v4.w};
                #pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const uint32_t v = vw[e];
                    const uint32_t vd = dup_lo(v);
                    uint32_t acc[NP];
                    #pragma unroll
                    for (int p = 0; p < NP; ++p) acc[p] = v + c[p][0];
                    #pragma unroll
                    for (int q = 1; q < D; ++q)
                        #pragma unroll
                        for (int p = 0; p < NP; ++p) acc[p] = v * acc[p] + c[p][q];
                    #pragma unroll
                    for (int q = 0; q + 1 < NH; q += 2) {
                        if ((q >> 1) & 1) a1 = and3(hne2(vd, h[q]), hne2(vd, h[q + 1]), a1);
                        else              a0 = and3(hne2(vd, h[q]), hne2(vd, h[q + 1]), a0);
                    }
                    if (NH & 1) a0 &= hne2(vd, h[NH - 1]);
                    #pragma unroll
                    for (int p = 0; p < NP; ++p) { if (p & 1) p3 |= (acc[p] == 0); else p2 |= (acc[p] == 0); }
                }
            }
            unsigned mask = __ballot_sync(0xFFFFFFFFu, ((a0 & a1) != 0xFFFFFFFFu) | p2 | p3);
            if (mask) found += __popc(mask);
        }
    }
    if (found) out[0] = found;
}

template <typename K>
void run(const char* name, K kern, const uint32_t* in, uint32_t* out, double rows, int sm, int ctas_per_sm) {
    const int grid = sm * ctas_per_sm, reps = 64;
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(a);
        kern<<<grid, 128>>>(in, out, reps);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms = 0; cudaEventElapsedTime(&ms, a, b);
        double rate = (double)grid * 128.0 * reps * TS * rows / (ms * 1e-3);
        if (rep) best = rate > best ? rate : best;
    }
    cudaError_t e = cudaGetLastError();
    printf("%-44s ctas/sm %2d  %.4e cells/s  (%.1f per clk per SM at 1.965 GHz)  %s\n", name, ctas_per_sm, best, best / sm / 1.965e9,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int sm = p.multiProcessorCount;
    printf("%s, %d SMs, clock %d kHz\n", p.name, sm, p.clockRate);
    static uint32_t h[4096];
    uint32_t x = 12345;
    for (int i = 0; i < 4096; ++i) { x = x * 1664525u + 1013904223u; h[i] = (x >> 2) & 0x3FFFFFFFu; }
    uint32_t *in, *out; cudaMalloc(&in, sizeof(h)); cudaMalloc(&out, 64); cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int c : {4}) {
        run("isetp R=16", k_isetp<16>, in, out, 16, sm, c);
        run("isetp R=32", k_isetp<32>, in, out, 32, sm, c);
        run("hset2 NH=8  (16 rows)", k_hset2<8>, in, out, 16, sm, c);
        run("hset2 NH=16 (32 rows)", k_hset2<16>, in, out, 32, sm, c);
        run("hset2 NH=24 (48 rows)", k_hset2<24>, in, out, 48, sm, c);
        run("mix hset2 NH=16 + isetp NI=4  (36 rows)", k_mix<16, 4>, in, out, 36, sm, c);
        run("mix hset2 NH=16 + isetp NI=8  (40 rows)", k_mix<16, 8>, in, out, 40, sm, c);
        run("mix hset2 NH=16 + isetp NI=12 (44 rows)", k_mix<16, 12>, in, out, 44, sm, c);
        run("mix hset2 NH=24 + isetp NI=8  (56 rows)", k_mix<24, 8>, in, out, 56, sm, c);
        run("horner NI=14 + 2 x deg 8 (30 rows)", k_horner<14, 2, 8>, in, out, 30, sm, c);
        run("horner NI=12 + 2 x deg 9 (30 rows)", k_horner<12, 2, 9>, in, out, 30, sm, c);
        run("horner NI=20 + 3 x deg 8 (44 rows)", k_horner<20, 3, 8>, in, out, 44, sm, c);
        run("horner-A NI=14 + 2 x deg 8 (30 rows)", k_horner_a<14, 2, 8>, in, out, 30, sm, c);
        run("horner-A NI=13 + 2 x deg 8 (29 rows)", k_horner_a<13, 2, 8>, in, out, 29, sm, c);
        run("hset2 NH=8 + 2 x deg 8 (32 rows)", k_hset2_horner<8, 2, 8>, in, out, 32, sm, c);
        run("hset2 NH=9 + 2 x deg 8 (34 rows)", k_hset2_horner<9, 2, 8>, in, out, 34, sm, c);
        run("hset2 NH=10 + 2 x deg 8 (36 rows)", k_hset2_horner<10, 2, 8>, in, out, 36, sm, c);
        run("hset2 NH=12 + 3 x deg 8 (48 rows)", k_hset2_horner<12, 3, 8>, in, out, 48, sm, c);
        run("horner-A NI=16 + 2 x deg 8 (32 rows)", k_horner_a<16, 2, 8>, in, out, 32, sm, c);
        run("horner-A NI=12 + 2 x deg 8 (28 rows)", k_horner_a<12, 2, 8>, in, out, 28, sm, c);
        run("horner-A NI=20 + 3 x deg 8 (44 rows)", k_horner_a<20, 3, 8>, in, out, 44, sm, c);
    }
    return 0;
}
