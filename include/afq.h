/* afq.h — C ABI of the B200-native `alevin-fry quant` hot path.
 *
 * The reference (alevin-fry 0.18.0, Rust) has no FFI seam; this header cuts one at the
 * "S-batch" boundary of SURVEY.md §8(b): it replaces the worker pool of
 * `run_worker_thread` (src/quant.rs:659-1325) — i.e. everything between "a MetaChunk of
 * collated per-cell chunks has been parsed" (src/quant.rs:733-735) and "the per-cell
 * sparse counts + featureDump statistics are known" (src/quant.rs:1150-1196, 1266-1268).
 *
 * Plain C: pointers and sizes only, no exceptions cross the boundary, every call returns
 * an int status (0 = AFQ_OK) and `afq_last_error` returns a message for the last failure.
 * The same structs are consumed by the CUDA product (libafq.so) and by the CPU oracle
 * (oracle/libafq_oracle.so, test infrastructure only — see oracle/README).
 */
#ifndef AFQ_H
#define AFQ_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AFQ_ABI_VERSION 3

/* status codes */
enum {
  AFQ_OK = 0,
  AFQ_ERR_INVALID = 1,     /* bad argument / inconsistent batch                         */
  AFQ_ERR_UNSUPPORTED = 2, /* valid request the CUDA path does not implement            */
  AFQ_ERR_CUDA = 3,        /* CUDA runtime error (message has the cudaError string)     */
  AFQ_ERR_NO_DEVICE = 4,   /* no usable sm_100 device — the product never falls back    */
  AFQ_ERR_INTERNAL = 5
};

/* `-r/--resolution`, numbered like ResolutionStrategy (src/quant.rs:81-91). */
enum {
  AFQ_RES_TRIVIAL = 0,
  AFQ_RES_CR_LIKE = 1,
  AFQ_RES_CR_LIKE_EM = 2,
  AFQ_RES_PARSIMONY_EM = 3,
  AFQ_RES_PARSIMONY = 4,
  AFQ_RES_PARSIMONY_GENE_EM = 5,
  AFQ_RES_PARSIMONY_GENE = 6
};

/* SplicedAmbiguityModel (src/quant.rs:62-67). Only winner-take-all is on the CUDA path. */
enum { AFQ_SA_WINNER_TAKE_ALL = 0, AFQ_SA_PREFER_AMBIG = 1 };

/* per-cell flag bits in afq_result.flags */
enum {
  AFQ_FLAG_TINY = 1,  /* took the tiny-cell fast path (src/quant.rs:794-846)            */
  AFQ_FLAG_ALT = 2,   /* a PUG component > large_graph_thresh was resolved cr-like
                         (src/pugutils.rs:1055-1072; quant.json alt_resolved_cell_numbers) */
  AFQ_FLAG_EMPTY = 4  /* no expressed gene (src/quant.rs:1173-1179)                      */
};

/* Mirrors WorkerConfig (src/quant.rs:398-416). */
typedef struct afq_config {
  int32_t resolution;         /* AFQ_RES_*                                              */
  int32_t usa_mode;           /* 3-column t2g: gene ids are 2k (spliced) / 2k+1 (unspl.) */
  int32_t em_init_uniform;    /* --init-uniform (EmInitType::Uniform) else Informative  */
  int32_t pug_exact_umi;      /* --umi-edit-dist 0                                      */
  int32_t sa_model;           /* AFQ_SA_* (prefer-ambig: src/pugutils.rs:505-641; applied as
                                 given — the host ignores it outside USA mode like
                                 src/quant.rs:1457-1467; switches the tiny-cell path off)   */
  int32_t dump_eq;            /* -d/--dump-eqclasses: keep every cell's gene eq-classes (afq_result_eqclasses)      */
  uint32_t num_gene_ids;      /* size of the tid_to_gid value space: G (gene mode) or 2G */
  uint32_t num_rows;          /* output columns: G (gene mode) or 3G (USA)               */
  uint64_t small_thresh;      /* --small-thresh (default 100); 0 disables the tiny path  */
  uint64_t large_graph_thresh;/* --large-graph-thresh (default 1000)                     */
  uint16_t barcode_len;       /* bases; informational                                    */
  uint16_t umi_len;           /* bases; <= 16 on the CUDA path (UMI packed in 32 bits)   */
  int32_t device;             /* CUDA device ordinal for this context (one ctx per GPU)  */
} afq_config;

/* One MetaChunk-like batch of consecutive collated cells, structure-of-arrays.
 * Replaces `Chunk<AlevinFryReadRecord>` (libradicl; used at src/quant.rs:470, 739-757).
 * All pointers are host pointers for afq_submit and device pointers for
 * afq_quant_device. Records of cell c are [cell_rec_offsets[c], cell_rec_offsets[c+1]).
 * `refs` hold transcript ids with the orientation bit already cleared, in the order the
 * mapper emitted them (ascending — src/eq_class.rs:859 relies on it).
 * Every record is expected to carry at least one alignment: the reference's RAD writer asserts it
 * (src/convert.rs:132) and its graph / EM resolvers index label[0], i.e. panic otherwise. cr-like and
 * trivial skip an alignment-free record here (as the reference's tiny-cell path does, src/quant.rs:495);
 * for the other resolutions the result of such a batch is unspecified — the callers that read files
 * (afqh_quantify) reject it as a corrupt collated RAD before it gets here.                          */
typedef struct afq_batch {
  uint64_t first_cell_index;        /* chunk index of cell 0 (src/quant.rs:734)          */
  uint64_t n_cells;
  uint64_t n_records;
  uint64_t n_refs_total;            /* < 2^32                                            */
  const uint64_t* cell_rec_offsets; /* [n_cells+1]                                       */
  const uint32_t* rec_umi32;        /* [n_records] 2-bit packed UMI (umi_len <= 16)      */
  const uint32_t* rec_ref_offsets;  /* [n_records+1] CSR into refs                       */
  const uint32_t* refs;             /* [n_refs_total]                                    */
  const uint8_t* rec_na8;           /* optional, [n_records]: alignment count per record
                                       (what the RAD record header stores, as one byte — every
                                       count must be <= 255) INSTEAD of rec_ref_offsets (pass
                                       NULL there): 1 byte instead of 4 per record over PCIe;
                                       the device derives the CSR offsets with a scan          */
  const uint8_t* rec_umi24;         /* optional, [3*n_records] little-endian 24-bit UMIs INSTEAD
                                       of rec_umi32 (pass NULL there); needs umi_len <= 12 (10x v3
                                       and every chemistry with a <= 12 bp UMI): 3 bytes instead
                                       of 4 per record over PCIe. 4-byte aligned pointer.       */
  const uint8_t* refs24;            /* optional, [3*n_refs_total] little-endian 24-bit transcript
                                       ids INSTEAD of refs (pass NULL there); needs every id
                                       < 2^24. 4-byte aligned pointer. The device widens both to
                                       u32 in HBM (k_unpack24) before the resolve kernels run.  */
} afq_batch;

/* Per-cell results in input cell order. CSR with ascending columns inside a row — the
 * order the reference's dense scan produces (src/quant.rs:1156-1168, 1266-1268).
 * Library-owned until afq_result_release (host results) / caller-owned (device API).   */
typedef struct afq_result {
  uint64_t n_cells;
  uint64_t nnz;
  const uint64_t* row_ptr;          /* [n_cells+1]                                       */
  const uint32_t* col;              /* [nnz]                                             */
  const float* val;                 /* [nnz]                                             */
  const float* sum_umi;             /* [n_cells] DeduplicatedReads (src/quant.rs:1162)   */
  const float* max_umi;             /* [n_cells]                                         */
  const uint32_t* num_expr;         /* [n_cells] NumGenesExpressed                       */
  const uint32_t* num_over_mean;    /* [n_cells] NumGenesOverMean (src/quant.rs:1192)    */
  const uint8_t* flags;             /* [n_cells] AFQ_FLAG_*                              */
} afq_result;

/* Caller-provided device buffers for afq_quant_device. Capacities in elements. The call
 * fails with AFQ_ERR_INVALID (nothing written past a capacity) if one is too small;
 * `col/val` never need more than n_refs_total entries.                                 */
typedef struct afq_device_out {
  uint64_t* row_ptr;    uint64_t cap_cells;   /* needs n_cells+1                         */
  uint32_t* col;        float* val;  uint64_t cap_nnz;
  float* sum_umi;       float* max_umi;
  uint32_t* num_expr;   uint32_t* num_over_mean;
  uint8_t* flags;
} afq_device_out;

typedef struct afq_ctx afq_ctx;

/* Create a context on cfg->device holding tid_to_gid (src/quant.rs:1422-1437, 1580) in
 * HBM. Fails with AFQ_ERR_NO_DEVICE when no CUDA device is usable.                     */
int afq_create(const afq_config* cfg, const uint32_t* tid_to_gid, uint64_t n_refs,
               afq_ctx** out);
void afq_destroy(afq_ctx* ctx);
const char* afq_last_error(const afq_ctx* ctx); /* ctx may be NULL: last create error    */

/* Host API (the e2e path): copies the batch H2D, runs the kernels, copies the result D2H.
 * afq_submit is asynchronous and may be called again before afq_wait (batches pipeline
 * over internal streams); tickets complete in submission order.                        */
int afq_submit(afq_ctx* ctx, const afq_batch* host_batch, uint64_t* ticket);
int afq_wait(afq_ctx* ctx, uint64_t ticket, afq_result* out);
void afq_result_release(afq_ctx* ctx, afq_result* res);

/* Device-resident API: every pointer in `dev_batch`/`out` is device memory on the
 * context's GPU; work is enqueued on `cuda_stream` (a cudaStream_t, NULL = default) and
 * the call returns without synchronising unless the batch needs the large-cell path.
 * `nnz_out` (host, may be NULL) receives nnz only if it forces a sync — prefer reading
 * row_ptr[n_cells] afterwards.                                                          */
int afq_quant_device(afq_ctx* ctx, const afq_batch* dev_batch, const afq_device_out* out,
                     void* cuda_stream);

/* Number of CUDA devices visible to the process (0 when there is none or the driver fails). */
int afq_device_count(void);

/* Pinned host memory for batches/results (cudaHostAlloc, portable across the contexts of
 * every GPU / cudaFreeHost).                                                            */
int afq_host_alloc(void** ptr, size_t bytes);
void afq_host_free(void* ptr);

/* Synchronise `cuda_stream`, report device-side error flags raised by the batches
 * enqueued on it (e.g. a cell larger than the giant-cell arena) and, if `nnz` and
 * `dev_row_ptr` are non-NULL, read back nnz = dev_row_ptr[n_cells].                     */
int afq_device_finish(afq_ctx* ctx, void* cuda_stream, uint64_t* nnz,
                      const uint64_t* dev_row_ptr, uint64_t n_cells);

/* --dump-eqclasses (src/quant.rs:1282-1307): with afq_config.dump_eq set, a host result also carries the gene-level
 * eq-classes (sorted gene-id label -> molecule count; the reference's per-cell `gene_eqc` map) of every cell that did
 * not take the tiny-cell fast path (those never build the map: src/quant.rs:1310-1321), classes in lexicographic label
 * order. Cell c owns classes [cell_cls_ptr[c], cell_cls_ptr[c+1]); class k has labels[cls_lab_ptr[k] .. cls_lab_ptr[k+1])
 * and counts[k]. Labels are gene ids as in tid_to_gid (USA: 2k / 2k+1). Valid until afq_result_release(res).      */
typedef struct afq_eqc_dump {
  uint64_t n_cells, n_classes, n_labels;
  const uint64_t* cell_cls_ptr;    /* [n_cells+1]   */
  const uint64_t* cls_lab_ptr;     /* [n_classes+1] */
  const uint32_t* labels;          /* [n_labels]    */
  const uint32_t* counts;          /* [n_classes]   */
} afq_eqc_dump;
int afq_result_eqclasses(afq_ctx* ctx, const afq_result* res, afq_eqc_dump* out);

/* `alevin-fry infer` (src/infer.rs:31-241): EM over a GLOBAL gene-eq-class table — the result of `quant
 * --dump-eqclasses` (gene_eqclass.txt.gz + geqc_counts.mtx). Replaces the worker loop around
 * em_optimize_subset_with_scratch (src/infer.rs:196-241, src/em.rs:251-456). All pointers are HOST pointers; the call
 * is synchronous. Rows of the count matrix are cells (CSR): cell c holds the classes cell_eq[cell_offsets[c] ..
 * cell_offsets[c+1]) with counts cell_cnt[..], in row order (the order the reference iterates them in). Labels are
 * column indices < num_rows (USA: the S | U | A column space written by --dump-eqclasses, src/quant.rs:273-353).
 * Uses the context's usa_mode / num_rows / em_init_uniform (infer itself always runs Informative init). The result
 * arrays are owned by the context and stay valid until the next afq_infer call or afq_destroy.            */
typedef struct afq_eqc_table {
  uint64_t n_classes;
  const uint32_t* label_offsets;   /* [n_classes+1] */
  const uint32_t* labels;          /* [label_offsets[n_classes]] */
} afq_eqc_table;
int afq_infer(afq_ctx* ctx, const afq_eqc_table* classes, uint64_t n_cells, const uint64_t* cell_offsets,
              const uint32_t* cell_eq, const uint32_t* cell_cnt, afq_result* out);

/* Introspection: ABI version; number of kernels this ctx has launched (bench.py's
 * gpu_launches); optional per-kernel device timing with CUDA events on the launching
 * stream (afq_set_profiling(1) -> run -> afq_profile_collect -> afq_profile_get(i)).    */
int afq_abi_version(void);
uint64_t afq_launch_count(const afq_ctx* ctx);
/* batches afq_wait ran a second time after growing a device arena (a giant cell, or the arena pools of the parsimony / EM
 * kernels, which afq_submit sizes from the batches seen so far instead of reading the device back) */
uint64_t afq_rerun_count(const afq_ctx* ctx);
int afq_set_profiling(afq_ctx* ctx, int enable);
int afq_profile_collect(afq_ctx* ctx);
int afq_profile_reset(afq_ctx* ctx);
int afq_profile_get(afq_ctx* ctx, int idx, const char** name, double* ms, uint64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* AFQ_H */
