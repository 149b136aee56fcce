/* mustache_b200 -- C ABI of the B200 scale-space engine.
 *
 * Drop-in boundary for the hot path of ay-lab/mustache v1.3.3.  The reference has no plugin API; the seam is the body
 * of `mustache()` between mustache/mustache.py:699 and :772 (and diff_mustache.py:262-425): given one dense N x N tile
 * of normalised contacts it produces, for every mask pixel, the best DoG response, its scale and its p-value.
 * Each entry point below cites the reference lines it replaces.  Plain C types only; every function returns 0 on
 * success or a negative MB200_ERR_* code and never throws or aborts; mb200_last_error() gives the message.
 *
 * Usage (one engine per GPU per process; an engine is not thread-safe, distinct engines are independent):
 *   mb200_create -> mb200_set_program -> mb200_configure -> mb200_upload_* (one per block) -> mb200_run
 *   -> mb200_block_counts / mb200_fetch_records -> ... -> mb200_destroy
 */
#ifndef MUSTACHE_B200_H
#define MUSTACHE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MB200_API __attribute__((visibility("default")))
#else
#define MB200_API
#endif

#define MB200_OK 0
#define MB200_ERR_CUDA (-1)       /* a CUDA call failed */
#define MB200_ERR_ARG (-2)        /* bad argument / call order */
#define MB200_ERR_CAPACITY (-3)   /* more records than the configured capacity; raise record_fraction and rerun */
#define MB200_ERR_NONFINITE (-4)  /* tile holds NaN/Inf (scipy.stats.expon.fit raises ValueError, mustache.py:755) */
#define MB200_ERR_NOMEM (-5)      /* device memory */

#define MB200_STEP_RESTART 1      /* chain cut before this Gaussian (first level of an octave that shares nothing) */
#define MB200_STEP_SCORE 2        /* score the previous DoG once this step's DoG exists (mustache.py:760-768) */
#define MB200_STEP_DIFFREF 4      /* this step's DoG is the octave's L_2 (diff_mustache.py:336, never rotated) */

typedef struct mb200_engine mb200_engine;

MB200_API int mb200_abi_version(void);
MB200_API int mb200_device_count(int* count);

/* One engine = one CUDA device + one stream + scratch.  Replaces the per-block multiprocessing.Process of
 * mustache.py:926-934 (regulator) as the unit of execution. */
MB200_API int mb200_create(int device, mb200_engine** out);
MB200_API void mb200_destroy(mb200_engine* e);
MB200_API const char* mb200_last_error(const mb200_engine* e);

/* Scale table: the chain of Gaussians with strictly increasing sigma (mustache.py:714-752 evaluated on the host with
 * numpy exactly as scipy does; see mustache_b200/ladder.py).  Per step: truncation radius, flags, and for scoring
 * steps the id reported in records (octave*12 + i, i.e. the reference's scales[o][i]).
 * half_taps: concatenated weights w[0..R] (centre first) of every step; tap_off[s] indexes into it. */
MB200_API int mb200_set_program(mb200_engine* e, int n_steps, const int32_t* radius, const int32_t* flags, const int32_t* score_id,
                      const int32_t* tap_off, const double* half_taps, int n_taps);

/* Geometry of the next batch: nblocks tiles of side n (CHUNK_SIZE, mustache.py:896), distance_in_px (mustache.py:892),
 * intra = chromosome == chromosome2 (mustache.py:705).  record_fraction: record capacity as a fraction of the band
 * pixels (<= 0 selects the default 1/8). */
MB200_API int mb200_configure(mb200_engine* e, int n, int dpx, int intra, int nblocks, double record_fraction);

/* Tile inputs, the state of `cc` at mustache.py:923-924 (normalised, scattered, NOT yet 2-filled).
 *  _coo_host   : block-local COO with unique coordinates (host pointers); entries off the band are ignored
 *  _dense_host : row-major n x n host tile, leading dimension ld (only diagonals 4..dpx+1 are transferred)
 *  _dense_dev  : same, device pointer (e.g. torch.Tensor.data_ptr())
 *  _band_host  : band layout [n][wsrc], element (i, k) = tile[i][i+4+k]  (host pointer) */
MB200_API int mb200_upload_coo_host(mb200_engine* e, int block, const int32_t* rows, const int32_t* cols, const double* vals,
                          int64_t nnz);
/* A run of blocks in one call: entries [offsets[b], offsets[b+1]) of the concatenated arrays are the block-local COO of
 * block first_block + b (offsets: nblk + 1 entries, offsets[0] = 0): one host -> device copy per array and one scatter
 * kernel for the whole batch instead of one of each per block.  Pageable arrays are staged before the call returns;
 * page-locked ones (mb200_host_alloc) are read by DMA and must stay valid until mb200_run / mb200_sync. */
MB200_API int mb200_upload_coo_batch(mb200_engine* e, int first_block, int nblk, const int64_t* offsets, const int32_t* rows,
                                     const int32_t* cols, const double* vals);
MB200_API int mb200_upload_coo_dev(mb200_engine* e, int block, const int32_t* rows_dev, const int32_t* cols_dev,
                                   const double* vals_dev, int64_t nnz);   /* same, device pointers, complete before the call */
MB200_API int mb200_upload_dense_host(mb200_engine* e, int block, const double* tile, int64_t ld);
MB200_API int mb200_upload_dense_dev(mb200_engine* e, int block, const double* tile_dev, int64_t ld);
MB200_API int mb200_upload_band_host(mb200_engine* e, int block, const double* band, int64_t wsrc);
/* Uploads are asynchronous on a second stream into the tile slot the kernels are NOT reading (tiles are double
 * buffered), so the uploads of batch k+1 overlap mb200_run of batch k (the analogue of the reference starting the next
 * `-p` processes, mustache.py:926-934).  Host buffers of the _dense_/_band_ uploads must stay valid until mb200_run /
 * mb200_sync; page-locked memory (mb200_host_alloc) makes them true DMA.  Results of batch k stay fetchable until the
 * next mb200_run. */

/* Runs the whole scale-space loop for all uploaded blocks: mask + fills (mustache.py:699-706), 12 Gaussians per octave
 * (:719-751), DoG (:728,738,754), 3x3 maxima (:740-743,757), extremum test and running best (:760-768), exponential fit
 * and p-value (:755-756).  Asynchronous on the engine's stream; mb200_sync / the fetch calls wait. */
MB200_API int mb200_run(mb200_engine* e);
MB200_API int mb200_sync(mb200_engine* e);
/* Two engines on one GPU, used alternately for a stream of batches (the reference keeps several blocks in flight with its
 * `-p` worker processes, mustache.py:926-934): everything enqueued on `e` from now on starts after the kernels `other` has
 * enqueued so far.  With it, the run of batch k+1 on one engine follows the run of batch k on the other back to back while
 * the post-processing and the result fetch of batch k (mb200_select_candidates ...) overlap it.  Same device required. */
MB200_API int mb200_run_after(mb200_engine* e, mb200_engine* other);

/* nz_count = np.sum(nz) (mustache.py:701; guards at :701 and :775 stay with the caller), n_found = sum(pAll != 2). */
MB200_API int mb200_block_counts(mb200_engine* e, int block, int64_t* nz_count, int64_t* n_found);

/* Records of every pixel with pAll != 2 (mustache.py:774), unordered: tile row, tile column, vAll, score id, raw p.
 * Caller-allocated arrays of `capacity` elements; *n_out receives the number of records the block has (may exceed
 * capacity, in which case only `capacity` are written). */
MB200_API int mb200_fetch_records(mb200_engine* e, int block, int64_t capacity, int32_t* rows, int32_t* cols, double* v,
                        int32_t* score_id, double* p, int64_t* n_out);

/* Whole-batch forms (the reference collects every block's loops in one Manager().list(), mustache.py:913-914, 959).
 * mb200_batch_counts: nz_count / n_found of every block (arrays of nblocks; no error for over-capacity blocks, so that a
 * caller can size a retry).  mb200_pack_records: packs the records of all blocks contiguously on the device, block b at
 * [offsets[b], offsets[b+1]) (offsets: nblocks + 1 entries); mb200_fetch_packed copies them to the host with one copy per
 * field (sigma / pair may be NULL); mb200_packed_device hands out the device arrays (for an NCCL gather without a host
 * round trip; valid until the next mb200_run / mb200_configure). */
MB200_API int mb200_batch_counts(mb200_engine* e, int64_t* nz_count, int64_t* n_found);
MB200_API int mb200_pack_records(mb200_engine* e, int64_t* offsets, int64_t* total);
MB200_API int mb200_fetch_packed(mb200_engine* e, int64_t capacity, int32_t* rows, int32_t* cols, double* v, int32_t* score_id,
                                 double* p, double* sigma, double* pair);
MB200_API int mb200_packed_device(mb200_engine* e, void** rows, void** cols, void** v, void** score_id, void** scored_index,
                                  void** p, void** sigma, void** pair);

/* Device half of the block post-processing, mustache.py:774-811 (and the per-map half of diff_mustache.py:428-500): for every
 * block of the batch, Benjamini-Hochberg over its found p-values (mustache.py:778; statsmodels' fdrcorrection formula, same
 * IEEE operations), selection `o < pt` (:791), the sparsity filter with numpy's slice semantics and threshold st (:800-811).
 * What leaves the GPU is one entry per SELECTED pixel: block, tile row / column, q (FDR), sigma (Scales), flags (bit 0: passed
 * the sparsity filter; bits 1-2: see mb200_enrich_candidates), cval (its value in the 2-filled tile, for the enrichment filter :822-828) and the 3 x 3 neighbourhoods
 * of the dense `o` and `so` matrices (:789-795; row-major, 1 off the mask, 2 / 1 on the mask but never updated), which is all
 * the clustering step (:830-848) reads.  candidate_fraction: capacity as a fraction of the batch's found records (<= 0: 1/16);
 * mb200_fetch_candidates returns MB200_ERR_CAPACITY when it was exceeded.  Order of the entries is unspecified.
 * mb200_fetch_q: q of every record of a block in the order of mb200_fetch_records (parity hook).
 * After mb200_run_differential (blocks 2k / 2k+1 = the two maps of pair k) three more neighbourhoods are available per
 * selected pixel, those the differential selection reads (diff_mustache.py:445-453, 567-568): `pair` and `v` of its own map
 * and `v` of the other map (1 off that map's mask, pPair / vAll where found, 2 / 0 on the mask but never updated); NULL skips. */
MB200_API int mb200_select_candidates(mb200_engine* e, double pt, double st, double candidate_fraction);
/* Enrichment filter of the selected candidates, mustache.py:816-828 (diff_mustache.py:511-527): c[x, y] > 2 * np.mean(non-zero
 * entries of the (y - x)-th diagonal of the 2-filled tile); np.mean's pairwise summation over the compacted diagonal is
 * reproduced bit for bit.  Optional, after mb200_select_candidates: sets flags bit 1 (passed) and bit 2 (decided) of every
 * candidate that passed the sparsity filter.  Synchronises (it sizes its scratch from the candidate count). */
MB200_API int mb200_enrich_candidates(mb200_engine* e);
MB200_API int mb200_fetch_candidates(mb200_engine* e, int64_t capacity, int32_t* block, int32_t* row, int32_t* col, int32_t* flags,
                                     double* q, double* sigma, double* cval, double* o9, double* so9, double* pair9,
                                     double* vself9, double* vother9, int64_t* n_out);
MB200_API int mb200_fetch_q(mb200_engine* e, int block, int64_t capacity, double* q, int64_t* n_out);
MB200_API int mb200_last_post_ms(mb200_engine* e, float* ms);

/* Arithmetic of the two Gaussian passes.  0 (default): the reference's own -- scipy's correlate1d multiplies, then adds
 * (two roundings per tap), which makes every Gaussian bit-identical to scipy.ndimage.gaussian_filter.  1: opt-in fast mode,
 * one fused multiply-add per tap (2R+1 instead of 3R+1 FP64 instructions per output); Gaussians then differ from scipy's in
 * the last bits (~1e-16 relative), which BASELINE's tolerance (coordinates and scale exact, p within 1e-6) absorbs on every
 * golden input (tests/test_gpu_fast_mode.py) but which is NOT the reference's arithmetic.  Takes effect at the next run. */
MB200_API int mb200_set_arithmetic(mb200_engine* e, int fused_multiply_add);

/* Kernel fusion, opt-in (0 = default: axis-0, axis-1 + DoG and scoring as three kernels).
 * 2: chains up to radius 14 (the default two octaves) run the axis-0 and the axis-1 pass as ONE kernel (kvh_kernel): the
 *    axis-0 results stay in shared memory instead of going through HBM, at 1.24 x the FP64 instructions.
 * 1: when the chain's widest axis-0 tile leaves room for three DoG levels in shared memory (the default two octaves do),
 * the axis-1 pass, the DoG and the scoring run as ONE kernel (khs_kernel) and the DoG levels never go to HBM; 0 (default):
 * always the three-kernel path, which measured 6 % faster on B200 (profiles/README.md: the fused kernel halves the DRAM
 * traffic but pays two CTA barriers per level).  Results are bit-identical either way (tests/test_gpu_configs.py).
 * Takes effect at the next mb200_configure. */
MB200_API int mb200_set_fusion(mb200_engine* e, int enable);

/* 1: batches of two or more blocks run as two half-batches on two streams (own halves of the scratch), so that the scoring
 * kernel of one half overlaps the Gaussian passes of the other; 0 (default): one stream.  Same results either way.  With
 * overlap the per-kernel times of mb200_last_timing overlap too (their sum exceeds the total).  Next mb200_run. */
MB200_API int mb200_set_overlap(mb200_engine* e, int enable);

/* Upper bound on the blocks one pass of the kernels handles (0 = as many as fit in device memory, the default).  The
 * reference's analogue is `-p`, the number of block processes alive at a time (mustache.py:931-934).  Takes effect at the
 * next mb200_configure. */
MB200_API int mb200_set_pass_limit(mb200_engine* e, int max_blocks);

/* Optional: the detection scale (sigma_i, mustache.py:767 `Scales[...] = scales[o][i]`) of every scoring step, in chain
 * order; call after mb200_set_program.  mb200_fetch_sigma then returns, per record and in the order of
 * mb200_fetch_records, the value the reference stores in `Scales` (0 when no table was registered). */
MB200_API int mb200_set_score_sigmas(mb200_engine* e, const double* sigma, int n_scored);
MB200_API int mb200_fetch_sigma(mb200_engine* e, int block, int64_t capacity, double* sigma, int64_t* n_out);

/* Device pointers of a block's record arrays (valid until the next mb200_run / mb200_configure; the engine's stream must
 * be synchronised -- mb200_block_counts does -- before they are read): lets a multi-GPU caller hand them to NCCL without a
 * host round trip.  scored_index is the 0-based index among the scoring steps (mb200_fetch_fits maps it to the score id). */
MB200_API int mb200_records_device(mb200_engine* e, int block, void** rows, void** cols, void** v, void** scored_index,
                                   void** p, int64_t* capacity);

/* Exponential fit per scored step of a block (loc = min|L|, scale = mean|L| - loc, mustache.py:755): arrays of n_scored. */
MB200_API int mb200_fetch_fits(mb200_engine* e, int block, double* loc, double* scale, int32_t* score_id, int capacity, int* n_scored);

/* Device time of the last mb200_run in milliseconds (CUDA events on the engine's stream):
 * prep (mask count), axis-0 kernel, axis-1 + DoG kernel, extremum/scoring kernel, statistics + p-values (after
 * mb200_run_differential: + the difference stack and pPair), total. */
MB200_API int mb200_last_timing(mb200_engine* e, float* prep_ms, float* kv_ms, float* kh_ms, float* ks_ms, float* fin_ms,
                                float* total_ms);
/* Kernel launches issued by the last mb200_run. */
MB200_API int mb200_last_launches(mb200_engine* e, int* launches);

/* Parity hooks: the Gaussian of chain step `step` and the DoG it completes, for block `block`, as dense n x n host
 * arrays (valid on the diagonals 2..dhi+2 the detector reads; zero elsewhere).  Either pointer may be NULL. */
MB200_API int mb200_debug_level(mb200_engine* e, int block, int step, double* gauss_out, double* dog_out);

/* Differential mode (diff_mustache.py:260-425).  Blocks 2k and 2k+1 of the batch hold map 1 and map 2 of pair k.
 * mb200_set_diff_program: the difference stack's chain -- per octave the Gaussians of levels 2 and 3, the second
 * flagged MB200_STEP_DIFFREF (the reference only ever uses L_2 = G_2 - G_3 of c1 - c2, diff_mustache.py:336, 371-378;
 * its `Lc` is never rotated, :413-425).  mb200_run_differential = mb200_run on both maps + the difference stack +
 * norm.fit over the common mask + the two-sided normal p (pPair, :372-385, 412, 421) of every record.  Octave of a
 * record = score_id / 12.  mb200_fetch_pair returns pPair in the same order as mb200_fetch_records. */
MB200_API int mb200_set_diff_program(mb200_engine* e, int n_steps, const int32_t* radius, const int32_t* flags,
                                     const int32_t* tap_off, const double* half_taps, int n_taps);
MB200_API int mb200_run_differential(mb200_engine* e);
MB200_API int mb200_fetch_pair(mb200_engine* e, int block, int64_t capacity, double* pair, int64_t* n_out);

/* The producer of the tile values: per-diagonal z-score normalisation of one chromosome's contact list, the reference's
 * normalize_sparse(x, y, v, resolution, distance_in_px) (mustache.py:622-686; both its windowed and its global branch).
 * x, y: bin indices, v: bias-corrected counts, normalised in place (host pointers, nnz entries).  weights (capacity
 * weights_cap, may be NULL) receives the reference's return value pval_weights, *n_weights its length.  np.mean / np.std
 * are reproduced bit for bit (numpy's pairwise summation); the 2 Mb box sums are taken left to right, which differs in
 * the last bits from the BLAS dot product behind np.convolve (whose order depends on the host CPU). */
MB200_API int mb200_normalize_sparse(mb200_engine* e, const int32_t* x, const int32_t* y, double* v, int64_t nnz,
                                     int resolution, int distance_in_px, double* weights, int weights_cap, int* n_weights);

/* Host-only inspection of the axis-0 kernel's plan (no engine, no device): the group every chain step is assigned to
 * (steps of a group share the folded pair sums x[-j] + x[j], which do not depend on the Gaussian) and the FP64
 * instructions per output pixel the plan costs, sum over groups of R_group*(2n+1)+n.  Either output may be NULL. */
MB200_API int mb200_kv_plan(int n_steps, const int32_t* radius, int32_t* group_of_step, int64_t* fp64_per_output);
/* Host-only inspection of the axis-1 kernel's staging ring (no engine, no device): for every chain step the offset and size
 * (in doubles) of its TMA box in the shared-memory ring and the latest earlier step whose box it overwrites (-1: none) --
 * the load of step s can start once step dep[s] has been released by every warp; ring_doubles receives the capacity.
 * Any output may be NULL. */
MB200_API int mb200_kh_ring_plan(int n_steps, const int32_t* radius, int32_t* offset, int32_t* size, int32_t* dep,
                                 int32_t* ring_doubles);

/* Native reader of the text contact format, the parse half of read_pd() (mustache.py:254-263: get_sep, pd.read_csv(header=None),
 * dropna, the is_chr filter on both chromosome columns of a 5-column file).  Host only, multi-threaded over a memory map
 * (threads <= 0: all cores).  Returns the rows of `chromosome` (all rows for a 3-column file) as the two position columns and
 * the value column; value_is_int tells whether pandas would have inferred int64 for it (every token an integer).  Strict:
 * MB200_PARSE_UNSUPPORTED (1) for anything pandas treats specially (quotes, ragged or missing fields, non-integer positions,
 * NaN tokens, values with more than 15 significant digits), in which case the caller stays with its pandas reader. */
#define MB200_PARSE_UNSUPPORTED 1
MB200_API int mb200_contacts_open(const char* path, const char* chromosome, int threads, void** handle, int64_t* n_rows, int* n_cols,
                                  int* value_is_int);
MB200_API int mb200_contacts_read(void* handle, int64_t* pos1, int64_t* pos2, double* value);
MB200_API void mb200_contacts_close(void* handle);

/* Pinned host memory for callers that want full-speed uploads. */
MB200_API int mb200_host_alloc(void** ptr, int64_t bytes);
MB200_API int mb200_host_free(void* ptr);

/* One-call drop-in for the scale-space half of `mustache(c, ...)` (mustache.py:697-772) on a host tile. */
MB200_API int mb200_scale_space_dense(mb200_engine* e, const double* tile, int n, int64_t ld, int dpx, int intra, int64_t capacity,
                            int32_t* rows, int32_t* cols, double* v, int32_t* score_id, double* p, int64_t* nz_count,
                            int64_t* n_found);

#ifdef __cplusplus
}
#endif
#endif
