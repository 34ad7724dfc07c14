/* afq_host.h — host-side drop-in for `alevin_fry::quant::quantify(QuantOpts)`
 * (reference src/quant.rs:359, src/prog_opts.rs:24-43) and the `alevin-fry quant` CLI
 * (src/main.rs:294-348, 633-821): reads <input-dir>/{generate_permit_list.json,
 * collate.json, map.collated.rad | map.collated.rad.sz}, the t2g TSV, and writes <output-dir>/{alevin/quants_mat.mtx,
 * alevin/quants_mat_rows.txt, alevin/quants_mat_cols.txt, featureDump.txt, quant.json}
 * (src/quant.rs:1588-1613, 1786-1847, 1913-1933). All per-cell compute goes through the
 * CUDA C-ABI of afq.h (afq_submit / afq_wait); there is no CPU compute path here.
 */
#ifndef AFQ_HOST_H
#define AFQ_HOST_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct afqh_quant_opts {          /* mirrors QuantOpts (src/prog_opts.rs:24-43)       */
  const char* input_dir;                  /* -i/--input-dir                                   */
  const char* tg_map;                     /* -m/--tg-map                                      */
  const char* output_dir;                 /* -o/--output-dir                                  */
  uint32_t num_threads;                   /* -t/--threads (host parse/format threads; floor 2) */
  uint32_t num_bootstraps;                /* -b (must be 0: bootstrap is not on the CUDA path) */
  int32_t init_uniform, summary_stat, dump_eq;
  const char* resolution;                 /* -r, case-insensitive                             */
  int32_t pug_exact_umi;                  /* derived from --umi-edit-dist                     */
  const char* sa_model;                   /* --sa-model (winner-take-all only)                */
  uint64_t small_thresh;                  /* --small-thresh                                   */
  uint64_t large_graph_thresh;            /* --large-graph-thresh                             */
  const char* filter_list;                /* --quant-subset or NULL                           */
  const char* cmdline;
  const char* version;
  int32_t device;                         /* CUDA device ordinal (used when `devices` is NULL) */
  uint64_t batch_records;                 /* records per device batch (0 = default 16M)       */
  const char* devices;                    /* NULL, "all", or a comma list of CUDA ordinals ("0,1,2,3"): ONE reader feeds
                                             the device batches round-robin to one context per GPU and ONE matrix is
                                             written in chunk order — the reference's one-reader / N-workers / one-matrix
                                             shape (src/quant.rs:1567-1575, 1678-1784, 1811-1847)                        */
  int32_t process_exits;                  /* the caller's process ends right after this call (the CLI): the GPU contexts and
                                             the pinned buffers are left to the operating system instead of being torn down
                                             (0.4-0.5 s of cudaFree / cudaFreeHost on a 100 k-cell run)                 */
} afqh_quant_opts;

/* Runs the whole quant stage. Returns 0 on success; on failure writes a message to err.    */
int afqh_quantify(const afqh_quant_opts* opts, char* err, size_t errlen);

/* `alevin-fry infer` (src/infer.rs:31-424, CLI src/main.rs:350-365, 823-848): EM from the files `quant --dump-eqclasses`
 * writes. Reads <count_mat> (geqc_counts.mtx: cells x eq-classes, `real` or `integer` MatrixMarket), <eq_labels>
 * (gene_eqclass.txt.gz), quants_mat_rows.txt / quants_mat_cols.txt beside the count matrix; writes
 * <output_dir>/{quants_mat.mtx, quants_mat_rows.txt, quants_mat_cols.txt}. The per-cell EM runs in afq_infer.     */
typedef struct afqh_infer_opts {
  const char* count_mat;      /* -c/--count-mat   */
  const char* eq_labels;      /* -e/--eq-labels   */
  const char* output_dir;     /* -o/--output-dir  */
  int32_t usa_mode;           /* --usa            */
  const char* filter_list;    /* --quant-subset or NULL */
  uint32_t num_threads;       /* -t (host threads; floor 2) */
  int32_t device;
} afqh_infer_opts;
int afqh_infer(const afqh_infer_opts* opts, char* err, size_t errlen);

/* Test/bench helper: write a collated RAD directory (map.collated.rad, collate.json,
 * generate_permit_list.json) in the wire format of SURVEY.md §8(b) from SoA arrays:
 * one chunk per cell, read tags b:u32|u64 (by bc_len), u:u32|u64 (by umi_len), alignment tag
 * compressed_ori_refid:u32 (orientation bit set = forward).                                */
int afqh_write_collated_rad(const char* dir, uint64_t n_cells, const uint64_t* cell_rec_offsets,
                            const uint64_t* cell_barcodes, const uint32_t* rec_umi32,
                            const uint32_t* rec_ref_offsets, const uint32_t* refs,
                            const char* const* ref_names, uint64_t n_refs, uint16_t bc_len,
                            uint16_t umi_len, char* err, size_t errlen);

/* CPU-only probe of a collated RAD file (plain, not .sz): runs the SAME prelude / tag-section / record-layout
 * code the quantifier uses and walks every chunk and record. Used by the tests to check the reader against
 * files written byte by byte from the reference's own statements (src/convert.rs:92-144, 254, 280-383,
 * 472-491) instead of against this repo's writer. Checksums: h = (h ^ value) * 0x100000001B3 over the values
 * in file order, starting from 0xCBF29CE484222325 (barcode of every record, UMI of every record, reference id of
 * every alignment with the orientation bit cleared).                                          */
typedef struct afqh_rad_info {
  uint64_t n_refs, num_chunks, n_records, n_alignments;
  uint64_t sum_bc, sum_umi, sum_refs;
  uint32_t bc_len, umi_len;                 /* cblen / ulen file tags (0 if absent)             */
  uint32_t read_bytes, aln_bytes, bc_size, umi_size, bc_off, umi_off, refid_off;
  uint32_t n_file_tags, n_read_tags, n_aln_tags;
} afqh_rad_info;
int afqh_rad_summary(const char* rad_path, afqh_rad_info* info, char* err, size_t errlen);

/* CPU-only run of the two host stages of `quant` (chunk index + the product's parallel parser; the parallel text formatting on
 * a synthetic result of the same shape): wall seconds per stage, and FNV checksums (as afqh_rad_summary) of the PARSED arrays —
 * UMI and alignment count of every record, reference id of every alignment, in file order.       */
typedef struct afqh_stage_info {
  uint64_t n_cells, n_records, n_alignments, nnz, mtx_bytes, mtx_sum;
  uint64_t sum_umi, sum_refs, sum_na;
  double walk_s, parse_s, format_s, parse_warm_s;
  uint32_t threads, pack24;
} afqh_stage_info;
int afqh_host_stage_bench(const char* rad_path, uint32_t n_threads, uint32_t frac_every, afqh_stage_info* out, char* err, size_t errlen);

/* Snappy FRAMING format (map.collated.rad.sz of `collate --compress`; the reference reads it with
 * snap::read::FrameDecoder, src/quant.rs:373-395) -> plain bytes. *out is malloc'ed (free with
 * afqh_free). Returns 0 on success.                                                          */
int afqh_snappy_framed_decompress(const uint8_t* src, size_t n, uint8_t** out, size_t* out_len, uint32_t n_threads,
                                  char* err, size_t errlen);
void afqh_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
