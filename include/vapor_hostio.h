/*
 * vapor_hostio.h -- C-ABI of the native region extraction that feeds the scoring path.
 *
 * The reference shells out per SV: `samtools faidx ref chr:a-b` inside ref_seq_readin
 * (vapor_vali/Simple_function.pyx:1203-1217) and `samtools view bam chr:a-b` inside chop_pacbio_read_by_pos
 * (:339-354), then walks every record's CIGAR in Python (cigar2alignstart_by_pos, :309-337) and keeps at most
 * 20 reads (minimize_pacbio_read_list, :1091-1102).  These entry points answer the same queries in-process, for
 * many windows per call and on several host threads, with the reference's exact rules and quirks:
 *   - regions are 1-based inclusive, clipped to the contig like samtools;
 *   - `samtools view` semantics: every record whose alignment overlaps the region, in file order, no FLAG or
 *     MAPQ filter; BAM records whose real CIGAR sits in the CG:B,I tag (more than 65535 operations) get it back;
 *   - CIGAR walk: S, M, =, I advance the read; M, =, D advance the reference; X, N, H, P advance NOTHING; the walk
 *     stops after the first operation that carries the reference position past start-1;
 *   - a record is kept when POS <= start, missed bases <= flank/2 and the read runs past the window; it is cut to
 *     end - start - miss_bp bases; Python's slicing rules apply to negative offsets;
 *   - at most `max_reads` reads per window, smallest miss_bp first, file order inside one miss_bp.
 * All of this is host code (C++17 + zlib inside libvapor_b200.so); no device is touched.
 * Every function returns 0 on success, a negative VAPOR_E_* code otherwise (text: vapor_io_last_error()).
 */
#ifndef VAPOR_HOSTIO_H
#define VAPOR_HOSTIO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* vapor_io_last_error(void);            /* last error of the calling thread */

/* ---- FASTA: samtools faidx (ref_seq_readin, Simple_function.pyx:1203-1217) ---------------------------------- */
int vapor_io_fasta_open(const char* path, void** fasta);        /* reads <path>.fai, writing it first when missing */
int vapor_io_fasta_close(void* fasta);
/* Bases start..end (1-based, inclusive) of `chrom`, clipped to the contig; unknown contig / empty interval -> 0 bases.
 * Writes min(*len, cap) bytes to `out` and the full length to *len. */
int vapor_io_fasta_fetch(void* fasta, const char* chrom, int64_t start, int64_t end, char* out, int64_t cap, int64_t* len);
/* Many regions in one call, `threads` host threads: region i is chrom[i] (NUL-terminated names back to back in
 * `chroms`, chrom_off[i] = offset of name i), start[i]..end[i].  Results concatenated into a library-owned buffer:
 * *bytes / off[n+1] stay valid until the next fetch_many on this handle or close. */
int vapor_io_fasta_fetch_many(void* fasta, int64_t n, const char* chroms, const int64_t* chrom_off, const int64_t* start,
                              const int64_t* end, int threads, const uint8_t** bytes, const int64_t** off);

/* ---- alignments: samtools view + chop_pacbio_read_by_pos (Simple_function.pyx:339-354) ----------------------- */
int vapor_io_aln_open(const char* path, void** aln);            /* SAM text (.sam, .sam.gz) or BAM (+ .bai when present) */
int vapor_io_aln_close(void* aln);

typedef struct vapor_io_reads {      /* reads of n windows; arrays owned by the result object */
    int64_t        n_win;
    const int64_t* win_off;          /* [n_win+1] reads of window w are [win_off[w], win_off[w+1])            */
    const int64_t* seq_off;          /* [n_reads+1] byte offsets into seq_bytes                              */
    const uint8_t* seq_bytes;        /* the cut reads, back to back (x[0] of the reference's [read, miss, qname]) */
    const int32_t* miss;             /* [n_reads] miss_bp (x[1])                                               */
    const int64_t* qname_off;        /* [n_reads+1] byte offsets into qname_bytes                              */
    const uint8_t* qname_bytes;      /* read names (x[2])                                                      */
    int64_t        n_records_seen;   /* records `samtools view` would have printed for all windows             */
} vapor_io_reads_t;

/* For every window w: the list chop_pacbio_read_by_pos(file, chrom[w], start[w], end[w], flank[w]) returns, over the
 * files alns[0..n_aln) in that order (bam_in_decide may name several files, Simple_function.pyx:69-89), then cut to
 * max_reads by minimize_pacbio_read_list when max_reads > 0.  *result must be released with vapor_io_reads_free. */
int vapor_io_chop_many(void* const* alns, int n_aln, int64_t n_win, const char* chroms, const int64_t* chrom_off,
                       const int64_t* start, const int64_t* end, const int64_t* flank, int max_reads, int threads,
                       vapor_io_reads_t** result);
int vapor_io_reads_free(vapor_io_reads_t* result);

/* Input-order gather of sharded results ("results written into a pre-sized host array indexed by input
 * position"): `src` holds n_runs runs of elements back to back, run r of run_len[r] elements goes to element run_dst[r]
 * of `dst`.  One run = the tasks of one SV.  Replaces the `cat` of the per-contig outputs in the reference's workflow
 * (wdl/VaPoRBedPerContig.wdl:102). */
int vapor_host_scatter_runs(void* dst, const void* src, int64_t elem_bytes, const int64_t* run_dst, const int64_t* run_len,
                            int64_t n_runs, int threads);

/* cigar2alignstart_by_pos (Simple_function.pyx:309-337) on a SAM CIGAR string: out[0] = read offset, out[1] = miss_bp. */
int vapor_io_cigar2alignstart(const char* cigar, int64_t align_start, int64_t start, int64_t* out);

#ifdef __cplusplus
}
#endif
#endif /* VAPOR_HOSTIO_H */
