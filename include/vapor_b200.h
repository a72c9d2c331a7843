/*
 * vapor_b200.h -- C-ABI of the B200-native VaPoR per-read scoring path.
 *
 * The reference (mills-lab/vapor) has no FFI: its "plugin API" is the Python
 * namespace `from vapor_vali.Simple_function import *` (vapor_vali/vapor:322,374,470)
 * and the hot path is called once per (read, SV) from the L2 drivers' read loops
 * (e.g. vapor_vali/Simple_function.pyx:1714-1726).  This library replaces the
 * body of those loops for a whole SV set in one call; the Python mirror in
 * vapor_b200/Simple_function.py keeps the reference signatures as thin wrappers.
 *
 * Each entry point cites the reference interface it replaces.  All pointers are
 * plain host pointers owned by the caller unless stated; no torch types cross
 * this boundary.  Every function returns 0 on success and a negative VAPOR_E_*
 * code on failure, with text available from vapor_gpu_last_error().
 * There is NO CPU fallback behind this API: without a CUDA device the calls fail.
 */
#ifndef VAPOR_B200_H
#define VAPOR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VAPOR_B200_ABI_VERSION 3

/* error codes */
#define VAPOR_OK            0
#define VAPOR_E_CUDA       -1   /* CUDA runtime error (text in last_error)          */
#define VAPOR_E_ARG        -2   /* malformed batch (bad index, k, mode, NULL)        */
#define VAPOR_E_CAPACITY   -3   /* a plot produced more hits than device memory holds */
#define VAPOR_E_STATE      -4   /* run/fetch without a resident batch                 */

/* scoring modes = which reference per-read driver a task stands for */
#define VAPOR_MODE_ABS          0 /* calcu_vapor_single_read_score_abs_dis_m1b                       Simple_function.pyx:182-203 */
#define VAPOR_MODE_W10          1 /* calcu_vapor_single_read_score_within_10Perc_m1b                 Simple_function.pyx:277-294 */
#define VAPOR_MODE_REDEF        2 /* calcu_vapor_single_read_score_directed_dis_m1b_redefine_diagnal Simple_function.pyx:241-257 */
#define VAPOR_MODE_ABS_AND_W10  3 /* simple-DEL rule: min of ABS and W10 scores                      Simple_function.pyx:1715-1726 */

/* per-task status */
#define VAPOR_ST_SKIPPED   0  /* `0 in pair`: the read contributes no score (e.g. Simple_function.pyx:1913) */
#define VAPOR_ST_SCORED    1
#define VAPOR_ST_BADREAD   2  /* read holds a character invert_base rejects: the reference raises KeyError (Simple_function.pyx:1421) */

/* Genotype codes: index into ['0/0','0/1','1/1'] (Simple_function.pyx:2062); 255 = 'NA' row (Simple_function.pyx:1231, 2087) */
#define VAPOR_GT_NA 255

/*
 * One batch = every (read, candidate-structure) pair of an SV set.
 * Sequences are the ASCII strings exactly as the reference sees them
 * (reads from chop_pacbio_read_by_pos, Simple_function.pyx:339-354; structures
 * from ref_seq_readin, Simple_function.pyx:1203-1217), concatenated CSR style.
 * Task i is one read of one SV/allele: x=[read, miss_bp, qname] (Simple_function.pyx:352)
 * scored against (ref_seq, alt_seq) with window_size k.
 */
typedef struct vapor_batch {
    const uint8_t* seq_bytes;      /* all sequences back to back                         */
    const int64_t* seq_off;        /* [n_seq+1] byte offsets into seq_bytes              */
    int64_t        n_seq;
    int64_t        n_task;
    const int32_t* task_read;      /* [n_task] sequence index of the read  (x[0])        */
    const int32_t* task_ref;       /* [n_task] sequence index of ref_seq                 */
    const int32_t* task_alt;       /* [n_task] sequence index of alt_seq                 */
    const int32_t* task_miss;      /* [n_task] miss_bp (x[1]); structures are cut [miss:] */
    const uint8_t* task_k;         /* [n_task] window_size, 1..40 (reference uses 10/20/30/40, Simple_function.pyx:2030-2046) */
    const uint8_t* task_mode;      /* [n_task] VAPOR_MODE_*                              */
    int64_t        n_sv;
    const int64_t* sv_task_off;    /* [n_sv+1] tasks of SV s are [off[s], off[s+1])      */
} vapor_batch_t;

/*
 * Results, caller-allocated.  Any pointer may be NULL to skip that output.
 * task_stat: the [a, b] pair the reference's calcu_* returns.  Columns 0,1 hold the
 * ABS / W10 / REDEF pair of the task's mode; for ABS_AND_W10 columns 0,1 = ABS pair
 * and 2,3 = W10 pair (else 2,3 = 0).
 * task_hits: len(ref_dotdata), len(alt_dotdata) for the first evaluation (cols 0,1)
 * and for the W10 evaluation of ABS_AND_W10 (cols 2,3).
 * task_hitsum: order-independent 64-bit checksum of each of those hit lists:
 *   sum over hits (with multiplicity) of vapor_hit_mix(x, y) mod 2^64 -- pins the
 *   matched-cell coordinates bit-exactly at any batch size.
 */
typedef struct vapor_out {
    double*   task_score;   /* [n_task]   per-read score (a16); valid when status==SCORED */
    uint8_t*  task_status;  /* [n_task]   VAPOR_ST_*                                       */
    double*   task_stat;    /* [n_task*4]                                                  */
    uint32_t* task_hits;    /* [n_task*4]                                                  */
    uint64_t* task_hitsum;  /* [n_task*4]                                                  */
    double*   sv_qs;        /* [n_sv] VaPoR_QS  result_organize_ins Simple_function.pyx:1226 */
    double*   sv_gs;        /* [n_sv] VaPoR_GS  Simple_function.pyx:1224                     */
    double*   sv_gq;        /* [n_sv] VaPoR_GQ  gt_estimate_log_likelihood Simple_function.pyx:2066 */
    uint8_t*  sv_gt;        /* [n_sv] VaPoR_GT code or VAPOR_GT_NA                          */
    int32_t*  sv_nscore;    /* [n_sv] number of reads that scored                           */
} vapor_out_t;

/* phase timings of the last run, milliseconds, CUDA events on the library's stream */
typedef struct vapor_timings {
    float h2d_ms, pack_ms, tile_ms, score_ms, genotype_ms, d2h_ms, total_ms;
    float host_prep_ms;          /* wall clock of the host-side operand/plot planning */
    int64_t cells;               /* recurrence cells evaluated (each distinct plot once) */
    int64_t hits;                /* hits emitted                                         */
    int64_t n_plots, n_operands, n_strips, n_waves, n_overflow_plots;
    int64_t launches;            /* kernels launched by the last run                     */
    int64_t bases;               /* bases packed by kernel 1                             */
    int64_t padded_cells;        /* cells the all-pairs tile kernel evaluates, strip padding included (k2_mode 0) */
    int64_t evaluated_cells;     /* word compares kernel 2 really made: padded_cells for the tile kernel; for the join
                                    kernel the sum over read k-mers of the size of the table bucket they fall into */
    float   table_ms;            /* kernel 1b: sorted word tables of the structure-side operands (k2_mode 1) */
    int32_t k2_mode;             /* kernel-2 variant the last plan was made for */
    int64_t table_bytes;         /* bytes of the sorted word tables kernel 1b writes and the join kernel stages (k2_mode 1) */
    int64_t probe_words;         /* read k-mer words the join kernel streams: sum of n over the non-empty plots */
    float   score_warp_ms;       /* the part of score_ms spent in the warp-per-task kernel 3 (k3w_score_reads) */
    int32_t pad_;
} vapor_timings_t;

/* Open one handle on CUDA device `device`.  One handle per device/thread; a handle is
 * not re-entrant, different handles may be driven from different threads/processes.
 * Replaces: nothing (the reference is a single CPU thread). */
int vapor_gpu_open(int device, void** handle);
int vapor_gpu_close(void* handle);
const char* vapor_gpu_last_error(void* handle);   /* handle may be NULL: last open() error */

/* Tunables: hit-buffer budget in bytes -- bounds device memory per wave (0 = default: a quarter of the memory free
 * at open(), at most 24 GB). */
int vapor_gpu_set_hit_budget(void* handle, int64_t bytes);
/* Named tunables (none changes a result): "hit_budget_bytes"; "k2_mode" (kernel 2: 1 = radix-partitioned join,
 * the default -- only cells whose k-mer words share a bucket are compared; 0 = all-pairs tile kernel -- every cell
 * of every plot is compared; both emit the identical hit set); "tile_variant" (inner loop of the tile
 * kernel: 0 = 16 rows/lane by ISETP only, 1/2/3/4 = 14/16/12/13 rows by ISETP + 2 row polynomials of
 * 8 rows by IMAD; default 4), "k2_ctas_per_sm" (persistent-grid size of the tile kernel), "k3_mode" (kernel 3:
 * 1 = one warp per task for the small value ranges, the default; 0 = one CTA per task everywhere;
 * "k3_warp_classes": how many of the scratch classes 2 048 / 4 096 / 8 192 / 16 384 / 26 624 bins go to the warp kernel,
 * default 3), "overlap" (1 = kernel 3 of a wave runs on a second stream under kernels 1-2 of the next wave; default 0),
 * "plan_threads" (host threads used for planning, 0 = auto).  Takes effect at the next upload. */
int vapor_gpu_set_option(void* handle, const char* name, int64_t value);

/* Blocking one-shot: host prep + H2D + kernels 1-4 + D2H.
 * Replaces the read loops `for x in all_reads: calcu_vapor_single_read_score_*(...)`
 * of every L2 driver (Simple_function.pyx:1520-1525, 1577-1581, 1630-1634, 1714-1726,
 * 1762-1766, 1815-1819, 1877-1882, 1911-1915, ...) plus result_organize_ins
 * (Simple_function.pyx:1219-1231) and gt_estimate_log_likelihood (Simple_function.pyx:2054-2069). */
int vapor_gpu_score(void* handle, const vapor_batch_t* in, vapor_out_t* out);

/* The same in three steps, so a caller can keep a batch resident in HBM:
 * upload = host prep + H2D, run = kernels only, fetch = D2H. */
int vapor_gpu_upload(void* handle, const vapor_batch_t* in);
int vapor_gpu_run(void* handle);
int vapor_gpu_fetch(void* handle, vapor_out_t* out);

int vapor_gpu_last_timings(void* handle, vapor_timings_t* t);

/* dotdata(kmerlen, seq1=read, seq2=structure) -> hits (Simple_function.pyx:545-549, 951-983).
 * Writes up to cap (x,y) pairs, sorted x ascending then y ascending with the
 * palindrome duplicates adjacent -- the reference's list order -- and the true
 * count to *n_hits (may exceed cap).  Returns VAPOR_E_ARG with status BADREAD text
 * when the read holds a character the reference's invert_base rejects. */
int vapor_gpu_dotdata(void* handle, int k, const uint8_t* read, int64_t read_len,
                      const uint8_t* structure, int64_t struct_len,
                      int32_t* xy, int64_t cap, int64_t* n_hits);

/* Self-plot quality control of window_size_refine (Simple_function.pyx:2030-2046): for every sequence i the
 * recurrence plot dotdata(k[i], seq_i, seq_i) is evaluated by kernels 1-2 and reduced on the fly to the
 * counts qual_check_repetitive_region (Simple_function.pyx:1154-1171) needs; no hit list is stored.
 * out[8*i ..]: len(dotdata), hits with x == y, hits with x > y, then min x, max x, min y, max y over the
 * x > y hits (0xFFFFFFFF / 0 when there is none), and VAPOR_ST_SCORED or VAPOR_ST_BADREAD (the sequence
 * holds a character invert_base rejects: the reference raises KeyError). */
int vapor_gpu_selfplot_qc(void* handle, const uint8_t* seq_bytes, const int64_t* seq_off, int64_t n_seq,
                          const uint8_t* k, int64_t* out);

/* Per-SV summary + genotype alone, for score lists that already sit on the host: kernel 4 on
 * scores[sv_off[s] .. sv_off[s+1]) for every SV s.  Replaces result_organize_ins (Simple_function.pyx:1219-1231)
 * and gt_estimate_log_likelihood (Simple_function.pyx:2054-2069) when the CLI calls them on a driver's
 * vapor_score_list (vapor_vali/vapor:339-340).  Outputs as in vapor_out_t. */
int vapor_gpu_summarize(void* handle, const double* scores, const int64_t* sv_off, int64_t n_sv,
                        double* sv_qs, double* sv_gs, double* sv_gq, uint8_t* sv_gt, int32_t* sv_nscore);

/* Pinned host memory for staging (optional; any host pointer is accepted by score/upload). */
int vapor_gpu_host_alloc(void** p, int64_t bytes);
int vapor_gpu_host_free(void* p);

/* Host-side planning alone (operands, plots, waves, kernel-2 and kernel-3 work lists) -- no device is touched.
 * For tests (the plan must not depend on the number of planning threads) and host-side tuning: *ms = wall time,
 * *digest = FNV-1a over every plan array the kernels read, counts[8] = operands, plots, tasks, waves, table chunks,
 * join items, recurrence cells, largest hit slab of a wave.  Any output pointer may be NULL. */
int vapor_host_plan(const vapor_batch_t* in, int k2_mode, int threads, int64_t wave_budget_bytes,
                    double* ms, uint64_t* digest, int64_t* counts);

/* Integer-issue microbenchmark used as the roofline denominator of the tile kernel:
 * which = 0: 32-bit compare-accumulate (ISETP) lane-ops/s, 1: LOP3 lane-ops/s, 2: IADD3 lane-ops/s (alu pipe alone),
 * 3: independent LOP3 + IMAD streams (alu pipe + fma pipe together: the dual-pipe integer issue rate the tile
 * kernel's inner loop is written for), 4: the tile kernel's inner loop in isolation (14 ISETP + 16 IMAD + 2 zero tests
 * per shared word, LDS.128 + one vote per 32 words, nothing else): integer lane-instructions/s,
 * 5: IMAD lane-ops/s (fma pipe alone), 6: independent ISETP + IMAD streams with the streamed operand first -- the tile
 * kernel's instruction mix without any of its structure (the non-circular dual-pipe ceiling). */
int vapor_gpu_int_peak(void* handle, int which, double* lane_ops_per_s);

/* The hit checksum mixer (host-callable; same function the kernels use). */
uint64_t vapor_hit_mix(uint32_t x, uint32_t y);

int vapor_b200_abi_version(void);
int vapor_gpu_device_count(void);     /* CUDA devices visible to this process (0 without a driver) */
/* PCI bus id of a device ("0000:1b:00.0", NUL-terminated) into buf[len]: lets a one-process-per-GPU launcher bind each
 * process to the NUMA node its GPU hangs off before it allocates pinned buffers (vapor_b200.engine.bind_to_gpu_numa). */
int vapor_gpu_pci_bus_id(int device, char* buf, int len);

#ifdef __cplusplus
}
#endif
#endif /* VAPOR_B200_H */
