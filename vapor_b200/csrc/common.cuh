// Shared definitions for the vapor_b200 kernels (sm_100a).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace vb {

// ---- alphabet after key_modify (reference: vapor_vali/Simple_function.pyx:908-949) ----
// codes: A0 C1 G2 T3 N4  a8 c9 g10 t11 n12, IUPAC RYSWKMBDHV -> N(4) / n(12), anything else 15.
constexpr int CODE_INVALID = 15;

// ---- k-mer words -------------------------------------------------------------------------
// bit 31: the word is a hash, not injective -- a match must be confirmed on the code strings
// bit 30: the k-mer is its own reverse complement (the reference appends such a dot twice)
// bits 0..29: exact canonical code (mixed) or hash.  No valid word is >= 0xFFFFFFFC.
constexpr uint32_t H_NEEDS_VERIFY   = 0x80000000u;
constexpr uint32_t H_PALINDROME     = 0x40000000u;
constexpr uint32_t H_STRUCT_INVALID = 0xFFFFFFFFu;  // structure k-mer that can never match (holds 'X', ...)
constexpr uint32_t H_READ_PAD       = 0xFFFFFFFEu;  // rejected read k-mer (the read's status is BADREAD)
constexpr uint32_t H_STREAM_PAD     = 0xFFFFFFFDu;  // tile kernel: padding of the streamed (shared-memory) axis
constexpr uint32_t H_ROW_PAD        = 0xFFFFFFFCu;  // tile kernel: padding of the register (row) axis
constexpr uint32_t H_MAX_VALID      = 0xFFFFFFFBu;

// operand flags
constexpr int OPF_UPPER = 1;   // str.upper() applied (ABS mode upper-cases ref/alt, Simple_function.pyx:183-184)
constexpr int OPF_READ  = 2;   // read side: reverse complement also matches; invalid characters are an error
constexpr int OPF_TABLE = 4;   // used as the structure side of a plot: the join kernel needs its sorted word table

struct Operand {           // one (sequence, k, casing, role): owns a k-mer hash array and a code array
    int64_t seq_begin;     // byte offset in seq_bytes
    int64_t hash_off;      // element offset in d_hash (multiple of 4)
    int64_t code_off;      // byte offset in d_code
    int32_t len;           // bases
    int32_t n;             // k-mers = max(0, len-k+1)
    int32_t k;
    int32_t flags;
};

struct Plot {              // one recurrence plot: read operand x structure operand cut at miss
    int64_t hit_off;       // element offset (uint2) of this plot's hit list
    int32_t read_op, struct_op;
    int32_t miss;          // structure is cut [miss:]
    int32_t n, m;          // read k-mers, structure k-mers after the cut
    uint32_t cap;          // capacity of the hit list
    int32_t kind;          // PLOT_QC bit: count into qc[hit_off*QC_WORDS ...] instead of appending hits at hit_off;
                           // PLOT_TAIL_T bit: the last (partial) read chunk is tiled transposed (see k2_tile.cuh)
    int32_t n_main_strips; // strips of the full read chunks; the tail strips follow
};
constexpr int PLOT_QC = 1, PLOT_TAIL_T = 2;

// Self-plot quality-control counters of one PLOT_QC plot (window_size_refine / qual_check_repetitive_region,
// Simple_function.pyx:2030-2046, 1154-1171): all hits, hits on the diagonal, hits below it (x > y) and
// the bounding box of those.
constexpr int QC_WORDS = 8;      // H, diag, lower, min x, max x, min y, max y (of the lower hits), spare

struct Task {              // one read of one SV/allele
    int32_t plot[4];       // ref, alt of evaluation A; ref, alt of the W10 evaluation of ABS_AND_W10 (else -1)
    int32_t len_ref, len_alt;   // full structure lengths: gate denominators (Simple_function.pyx:188-189)
    int32_t read_op;
    int32_t mode;
};

// ---- join variant of kernel 2 (k2_join.cuh) -----------------------------------------------------
// A structure-side operand is cut into chunks of at most K2J_CH k-mer positions; kernel 1b turns every chunk into a
// *table*: its valid words counting-sorted by the top `bits` bits of the 30-bit payload, the position of every
// sorted word, the 2^bits + 1 bucket offsets, and a membership bitmap over the top `fbits` bits (2^fbits >= 8 len, 12..16:
// at most one bit in eight is set) -- the pre-filter: three quarters of the read words match nothing and leave after one
// shared-memory load.
// One blob per chunk, staged into shared memory by one TMA bulk copy:
//   uint32 word[Lp] | uint16 pos[Lp] | uint16 off[2^bits + 8] | uint32 filter[2^fbits / 32]      Lp = len rounded up to 8
#ifndef K2J_CH_N
#define K2J_CH_N 8192
#endif
constexpr int K2J_CH       = K2J_CH_N;         // positions per table chunk (pos fits 16 bits; blob <= 66 KB)
constexpr int K2J_MIN_BITS = 8;
#ifndef K2J_MAX_BITS_N
#define K2J_MAX_BITS_N 12
#endif
constexpr int K2J_MAX_BITS = K2J_MAX_BITS_N;

struct TabChunk {
    int64_t blob_off;      // byte offset of the blob in d_table (16-byte aligned)
    int32_t op;            // operand the chunk belongs to
    int32_t pos0;          // first k-mer position covered
    int32_t len;           // positions covered (1 .. K2J_CH)
    int32_t bits;          // bucket key bits
    int32_t blob_bytes;    // multiple of 16
    int32_t pad_;
};
__host__ __device__ inline int k2j_lp(int len) { return (len + 7) & ~7; }
__host__ __device__ inline int k2j_bits(int len) {
    int b = K2J_MIN_BITS;
    while (b < K2J_MAX_BITS && (1 << b) < len) ++b;
    return b;
}
__host__ __device__ inline int k2j_fbits(int len) {
    int f = 12;
    while (f < 16 && (1 << f) < 8 * len) ++f;
    return f;
}
__host__ __device__ inline int k2j_blob_bytes(int len, int bits) {
    return 6 * k2j_lp(len) + 2 * ((1 << bits) + 8) + (1 << k2j_fbits(len)) / 8;
}

struct JoinItem {          // one CTA of the join kernel: one table chunk against a few plots that use it
    int32_t chunk;
    int32_t jp_begin, jp_end;   // range in the launch's plot-id list
    int32_t pad_;
};

// hits carry a flag bit in y while kernel 3 works on them (REDEF statistics)
constexpr uint32_t HIT_Y_MASK   = 0x0FFFFFFFu;
constexpr uint32_t HIT_F_CLEAN  = 0x80000000u;  // member of the cleaned dot set

__host__ __device__ __forceinline__ uint64_t hit_mix(uint32_t x, uint32_t y) {
    uint64_t z = ((uint64_t)x << 32) | (uint64_t)y;
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ int comp_code(int c) {
    // invert_base in code space (Simple_function.pyx:20): A<->T, C<->G, N->N, case preserved
    return (c & 4) ? c : (c ^ 3);
}

}  // namespace vb
