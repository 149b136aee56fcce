// C-ABI host side of the scale-space engine (see include/mustache_b200.h for the contract and the reference
// lines each entry point replaces).  Owns one CUDA stream and grow-only device scratch; no torch types anywhere.
#include "../../include/mustache_b200.h"
#include "mb_kernels.cuh"
#include "mb_normalize.cuh"
#include "mb_sort.cuh"
#include "mb_post.cuh"
#include "mb_parse.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
};

}  // namespace

struct mb200_engine {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t up_stream = nullptr;     // uploads run here, into the tile slot the compute stream is not reading
    cudaEvent_t ev_up = nullptr;          // uploads of the pending batch are complete
    cudaEvent_t ev_run[2] = {nullptr, nullptr};   // the run that read slot k is complete
    int slot_up = 0;                      // tile slot the next uploads go to
    int slot_run = 0;                     // tile slot of the last run
    bool slot_used[2] = {false, false};   // a run has read this slot (uploads into it must wait for ev_run)
    bool up_dirty = false;                // uploads happened since the last run
    char err[512] = {0};
    MbProgram prog;
    MbProgram dprog;                 // difference-stack chain (diff_mustache): G_2, G_3 of every octave
    KvPlan kvplan, dkvplan;          // kv_kernel's grouping of the two chains
    KvPlan pairplan;                 // kvh_kernel: the main chain in consecutive pairs
    DevBuf d_pairplan;
    bool kvh_smem_set = false;
    bool have_prog = false;
    bool have_dprog = false;
    bool ran_diff = false;
    bool configured = false;
    bool ran = false;
    int n = 0, dpx = 0, intra = 1, dhi = 0, wc = 0, vlo = 0, wv = 0, wl = 0, nblocks = 0, pass_blocks = 0;
    long long rec_cap = 0;
    int ncta_h = 0;
    int npass_run = 1;                // passes of the last mb200_run
    DevBuf raw, V, Lb, part_min, part_sum, rec_count, nz_count, nonfinite, rec_row, rec_col, rec_v, rec_sidx, rec_p,
        fit_loc, fit_scale, st_rows, st_cols, st_vals, st_dense, dbgG, dbgL, rawD, dout, dmu, dsd, rec_pair, d_score_id, d_score_sigma, rec_sid, rec_sigma,
        nz_xs, nz_ds, nz_perm, nz_vs, nz_out, nz_seg, nz_mean, nz_sd, nz_w, nz_lines, st_offsets,
        nz_x, nz_y, nz_v, sort_keys[2], sort_vals[2], sort_hist,
        rec_q, bh_tmin, slotmap, cd_block, cd_row, cd_col, cd_flags, cd_q, cd_sigma, cd_cval, cd_o9, cd_so9, cd_count, cd_slot, cd_pair9, cd_vs9, cd_vo9, dpart, en_need, en_lines, en_nlist, en_mean, en_scratch;
    bool post_diff = false;
    long long cand_cap = 0;
    bool post_done = false;
    cudaEvent_t ev_post0 = nullptr, ev_post1 = nullptr, ev_diff = nullptr;   // ev_diff: end of the difference-stack kernels
    float t_post = 0;
    std::vector<long long> h_offsets;
    std::vector<unsigned long long> h_nz, h_rec;
    std::vector<int> h_nonfinite;
    bool counts_valid = false;
    cudaEvent_t ev_begin = nullptr, ev_prep = nullptr, ev_end = nullptr;
    std::vector<cudaEvent_t> ev_pass;   // 4 per pass: start, after kv, after kh, after ks
    float t_prep = 0, t_kv = 0, t_kh = 0, t_ks = 0, t_fin = 0, t_total = 0;
    int launches = 0;
    size_t kv_smem_set = 0, kh_smem_set = 0, ks_smem_set = 0;
    bool kf_smem_set = false;
    double score_sigma[MB_MAX_STEPS] = {0};   // detection scale per scored index (mb200_set_score_sigmas), 0 if unset
    MbTensorMaps tmaps;              // main chain (host copy)
    MbTensorMaps dtmaps;             // difference chain (V boxes follow its radii)
    DevBuf d_tmaps, d_dtmaps;        // device copies the kernels read the descriptors from
    long long plane_v = 0, plane_l = 0;
    int overlap = 0;                 // mb200_set_overlap: 1 = two passes in flight on two streams (scoring of one half of the
                                     // batch overlaps the Gaussian passes of the other half)
    cudaStream_t stream2 = nullptr;  // second compute stream of the overlapped mode
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaStream_t post_stream = nullptr;   // high priority: BH / selection / filters / candidate fetch of a finished run; next to a
                                          // run of ANOTHER engine (mb200_run_after) its small grids are not queued behind that
                                          // run's large ones
    cudaEvent_t ev_post_done = nullptr;   // last work queued on post_stream; the next run on `stream` waits for it
    cudaEvent_t ev_chain = nullptr;  // mb200_run_after: marks what this engine has enqueued when another engine chains behind it
    int fusion = 0;                  // mb200_set_fusion: 1 = axis-1 + scoring fused (khs_kernel) whenever the chain fits, 0 = never
    int fast = 0;                    // mb200_set_arithmetic: 0 = the reference's multiply-then-add, 1 = fused multiply-add
    int pass_limit = 0;              // mb200_set_pass_limit: upper bound on blocks per pass (0 = as many as fit)
    int ndiff = 0;                   // MB_FLAG_DIFFREF steps of the difference chain
    // packed view of the batch's records (mb200_pack_records): block b occupies [pk_off[b], pk_off[b+1])
    DevBuf pk_row, pk_col, pk_v, pk_sid, pk_p, pk_sigma, pk_pair, pk_sidx, pk_offsets;
    std::vector<long long> pk_off;
    bool packed = false;
};

namespace {

int fail(mb200_engine* e, int code, const char* fmt, ...) {
    if (e) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(e->err, sizeof(e->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

#define CU(e, call)                                                                                           \
    do {                                                                                                      \
        cudaError_t _st = (call);                                                                             \
        if (_st != cudaSuccess)                                                                               \
            return fail((e), _st == cudaErrorMemoryAllocation ? MB200_ERR_NOMEM : MB200_ERR_CUDA, "%s: %s",   \
                        #call, cudaGetErrorString(_st));                                                      \
    } while (0)

int ensure(mb200_engine* e, DevBuf& b, size_t bytes) {
    if (bytes <= b.cap) return MB200_OK;
    if (b.p) {
        CU(e, cudaStreamSynchronize(e->stream));
        if (e->stream2) CU(e, cudaStreamSynchronize(e->stream2));
        if (e->post_stream) CU(e, cudaStreamSynchronize(e->post_stream));
        CU(e, cudaStreamSynchronize(e->up_stream));
        CU(e, cudaFree(b.p));
        b.p = nullptr;
        b.cap = 0;
    }
    CU(e, cudaMalloc(&b.p, bytes));
    b.cap = bytes;
    return MB200_OK;
}

void release(DevBuf& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.cap = 0;
}

int use_device(mb200_engine* e) {
    CU(e, cudaSetDevice(e->device));
    return MB200_OK;
}

constexpr size_t V_GUARD_BYTES = 64 * 1024;   // slack on both sides of the axis-0 scratch for 16-byte-aligned row copies

// Tiles are double-buffered: uploads of the next batch (on up_stream) overlap the kernels of the current one.
double* raw_slot(const mb200_engine* e, int slot) {
    return (double*)e->raw.p + (size_t)slot * e->nblocks * e->n * e->wc;
}

// called by every upload before it touches the upload slot
int begin_upload(mb200_engine* e) {
    if (!e->up_dirty && e->slot_used[e->slot_up]) {
        CU(e, cudaStreamWaitEvent(e->up_stream, e->ev_run[e->slot_up], 0));
        e->slot_used[e->slot_up] = false;
    }
    e->up_dirty = true;
    return MB200_OK;
}

// the batch that was just uploaded becomes the one the kernels read; the other slot takes the next uploads
int adopt_uploads(mb200_engine* e) {
    if (e->up_dirty) {
        CU(e, cudaEventRecord(e->ev_up, e->up_stream));
        CU(e, cudaStreamWaitEvent(e->stream, e->ev_up, 0));
        e->slot_run = e->slot_up;
        e->slot_up ^= 1;
        e->up_dirty = false;
    }
    return MB200_OK;
}

size_t v_bytes_per_block(const mb200_engine* e) {
    return (size_t)e->prog.n_steps * e->plane_v * sizeof(double);
}

size_t l_bytes_per_block(const mb200_engine* e) {
    return (size_t)e->prog.n_steps * e->plane_l * sizeof(double);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// Skewed 3-D view of a band-layout scratch array (see MbTensorMaps): x = column index, y = image row, z = plane.
int encode_skewed(mb200_engine* e, CUtensorMap* map, void* base, int row_len, long long plane, int planes, int box_w) {
    EncodeTiledFn fn = encode_tiled();
    if (!fn) return fail(e, MB200_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[3] = {(cuuint64_t)(e->n + row_len), (cuuint64_t)e->n, (cuuint64_t)planes};
    const cuuint64_t strides[2] = {(cuuint64_t)(row_len - 1) * sizeof(double), (cuuint64_t)plane * sizeof(double)};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)KH_TR, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(e, MB200_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) row_len=%d box=%d", (int)r, row_len, box_w);
    return MB200_OK;
}

int encode_maps(mb200_engine* e, const MbProgram& pg, MbTensorMaps& tm, bool with_l) {
    void* vbase = (char*)e->V.p + V_GUARD_BYTES;
    void* lbase = (char*)e->Lb.p + V_GUARD_BYTES;
    const int planes_v = e->prog.n_steps * e->pass_blocks;       // capacity of the scratch arrays
    memset(&tm, 0, sizeof(tm));
    for (int s = 0; s < pg.n_steps; ++s) {
        int st = encode_skewed(e, &tm.v[s], vbase, e->wv, e->plane_v, planes_v, kh_box_width(pg.st[s].radius));
        if (st) return st;
    }
    if (with_l && pg.pad) {
        for (int s = 0; s < pg.n_steps; ++s) {
            int st = encode_skewed(e, &tm.vf[s], vbase, e->wv, e->plane_v, planes_v, kf_box_width(pg.st[s].radius, pg.pad));
            if (st) return st;
        }
    }
    if (with_l) return encode_skewed(e, &tm.l, lbase, e->wl, e->plane_l, planes_v, KS_PITCH);
    return MB200_OK;
}

MbGeom make_geom(mb200_engine* e, int first_block, int nblk) {
    MbGeom g;
    g.n = e->n;
    g.dpx = e->dpx;
    g.intra = e->intra;
    g.dhi = e->dhi;
    g.wc = e->wc;
    g.vlo = e->vlo;
    g.wv = e->wv;
    g.nblk = nblk;
    g.ncta_h = e->ncta_h;
    g.dbg_step = -1;
    g.rec_cap = e->rec_cap;
    const size_t ns = (size_t)std::max(e->prog.n_scored, 1);
    g.raw = raw_slot(e, e->slot_run) + (size_t)first_block * e->n * e->wc;
    g.V = (double*)((char*)e->V.p + V_GUARD_BYTES);     // bulk copies may start a few elements before a row
    g.L = (double*)((char*)e->Lb.p + V_GUARD_BYTES);
    g.wl = e->wl;
    g.plane_v = e->plane_v;
    g.plane_l = e->plane_l;
    g.part_min = (double*)e->part_min.p + (size_t)first_block * ns * e->ncta_h;
    g.part_sum = (double*)e->part_sum.p + (size_t)first_block * ns * e->ncta_h;
    g.rec_count = (unsigned long long*)e->rec_count.p + first_block;
    g.rec_row = (int*)e->rec_row.p + (size_t)first_block * e->rec_cap;
    g.rec_col = (int*)e->rec_col.p + (size_t)first_block * e->rec_cap;
    g.rec_v = (double*)e->rec_v.p + (size_t)first_block * e->rec_cap;
    g.rec_sidx = (int*)e->rec_sidx.p + (size_t)first_block * e->rec_cap;
    g.dbgG = nullptr;
    g.dbgL = nullptr;
    g.fill = 2.0;                      // mustache.py:703-706
    g.dout = nullptr;
    g.ndiff = e->ndiff;
    g.zstride = e->pass_blocks;
    g.zoff = 0;
    return g;
}

int kv_tile_rows(const mb200_engine* e) { return e->wv < 1024 ? KV_TH_NARROW : KV_TH_WIDE; }

dim3 kv_grid(const mb200_engine* e, int nblk) {
    const int th = kv_tile_rows(e);
    const int span = e->wv + th - 1;
    return dim3((span + KV_TW - 1) / KV_TW, (e->n + th - 1) / th, nblk);
}

dim3 kh_grid(const mb200_engine* e, int nblk) {
    const int span = (e->dhi + 1) + KH_TR - 1;            // diagonals 2..dhi+2 over the rows of one tile
    return dim3((span + KH_TC - 1) / KH_TC, (e->n + KH_TR - 1) / KH_TR, nblk);
}

dim3 ks_grid(const mb200_engine* e, int nblk) {
    const int span = (e->dhi - 4 + 1) + KS_SR - 1;
    return dim3((span + KS_SC - 1) / KS_SC, (e->n + KS_SR - 1) / KS_SR, nblk);
}

dim3 kf_grid(const mb200_engine* e, int nblk, int tc) {
    const int span = (e->dhi - 4 + 1) + KS_SR - 1;
    return dim3((span + (tc - 2) - 1) / (tc - 2), (e->n + KS_SR - 1) / KS_SR, nblk);
}

// tile columns of the fused kernel for this batch (0: three-kernel path)
int fused_tc(const mb200_engine* e) { return (e->fusion == 1 && e->prog.n_scored > 0) ? e->prog.pad : 0; }

// axis-0 + axis-1 in one kernel (kvh_kernel) for this batch?
bool fused_vh(const mb200_engine* e) { return e->fusion == 2 && e->pairplan.n_groups > 0; }

int set_smem_limits(mb200_engine* e) {
    const size_t kvb = kv_smem_bytes(e->prog.rmax, KV_TH_WIDE), khb = kh_smem_bytes(e->prog.rmax, e->prog.n_scored);
    if (kvb > 227 * 1024 || khb > 227 * 1024)
        return fail(e, MB200_ERR_ARG, "radius %d needs %zu / %zu bytes of shared memory (> 227 KB)", e->prog.rmax, kvb, khb);
    if (kvb != e->kv_smem_set) {
        CU(e, cudaFuncSetAttribute(kv_kernel<KV_TH_WIDE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kvb));
        CU(e, cudaFuncSetAttribute(kv_kernel<KV_TH_NARROW, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kvb));
        CU(e, cudaFuncSetAttribute(kv_kernel<KV_TH_WIDE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kvb));
        CU(e, cudaFuncSetAttribute(kv_kernel<KV_TH_NARROW, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kvb));
        CU(e, cudaFuncSetAttribute(kv_kernel<KV_TH_WIDE, false, KV_GSMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kvb));
        CU(e, cudaFuncSetAttribute(kv_kernel<KV_TH_NARROW, false, KV_GSMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kvb));
        CU(e, cudaFuncSetAttribute(kv_kernel<KV_TH_WIDE, true, KV_GSMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kvb));
        CU(e, cudaFuncSetAttribute(kv_kernel<KV_TH_NARROW, true, KV_GSMALL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kvb));
        e->kv_smem_set = kvb;
    }
    if (khb != e->kh_smem_set) {
        CU(e, cudaFuncSetAttribute(kh_kernel<KH_MAIN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)khb));
        CU(e, cudaFuncSetAttribute(kh_kernel<KH_DIFF, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)khb));
        CU(e, cudaFuncSetAttribute(kh_kernel<KH_DEBUG, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)khb));
        CU(e, cudaFuncSetAttribute(kh_kernel<KH_MAIN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)khb));
        CU(e, cudaFuncSetAttribute(kh_kernel<KH_DIFF, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)khb));
        CU(e, cudaFuncSetAttribute(kh_kernel<KH_DEBUG, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)khb));
        e->kh_smem_set = khb;
    }
    if (!e->kf_smem_set) {
        CU(e, cudaFuncSetAttribute(khs_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kf_smem_bytes(64)));
        CU(e, cudaFuncSetAttribute(khs_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kf_smem_bytes(64)));
        CU(e, cudaFuncSetAttribute(khs_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kf_smem_bytes(128)));
        CU(e, cudaFuncSetAttribute(khs_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kf_smem_bytes(128)));
        e->kf_smem_set = true;
    }
    if (!e->kvh_smem_set) {
        CU(e, cudaFuncSetAttribute(kvh_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kvh_smem_bytes()));
        CU(e, cudaFuncSetAttribute(kvh_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kvh_smem_bytes()));
        e->kvh_smem_set = true;
    }
    const size_t ksb = ks_smem_bytes(e->prog.n_scored);
    if (ksb != e->ks_smem_set) {
        CU(e, cudaFuncSetAttribute(ks_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ksb));
        e->ks_smem_set = ksb;
    }
    return MB200_OK;
}

int launch_pass(mb200_engine* e, int first_block, int nblk, MbGeom* dbg_geom, cudaEvent_t after_kv,
                const MbProgram* program = nullptr, cudaEvent_t after_kh = nullptr, cudaStream_t sq = nullptr, int zoff = 0) {
    MbGeom g = dbg_geom ? *dbg_geom : make_geom(e, first_block, nblk);
    if (!sq) sq = e->stream;
    g.zoff = zoff;
    const MbProgram& pg = program ? *program : e->prog;
    const int th = kv_tile_rows(e);
    const size_t kvb = kv_smem_bytes(pg.rmax, th), khb = kh_smem_bytes(pg.rmax, pg.n_scored);
    const MbTensorMaps* tm = (const MbTensorMaps*)(program ? e->d_dtmaps.p : e->d_tmaps.p);
    const KvPlan& kp = program ? e->dkvplan : e->kvplan;
    const dim3 gv = kv_grid(e, nblk), gh = kh_grid(e, nblk);
    const bool plain = program == nullptr && g.dout == nullptr && g.dbgG == nullptr && g.dbgL == nullptr;
    if (plain && fused_vh(e)) {
        // axis-0 and axis-1 pass in one kernel: the axis-0 results stay in shared memory; scoring as usual
        if (e->fast) kvh_kernel<true><<<gh, KH_THREADS, kvh_smem_bytes(), sq>>>(pg, (const KvPlan*)e->d_pairplan.p, g);
        else kvh_kernel<false><<<gh, KH_THREADS, kvh_smem_bytes(), sq>>>(pg, (const KvPlan*)e->d_pairplan.p, g);
        CU(e, cudaGetLastError());
        if (after_kv) CU(e, cudaEventRecord(after_kv, sq));
        if (after_kh) CU(e, cudaEventRecord(after_kh, sq));
        e->launches += 1;
        if (pg.n_scored > 0) {
            ks_kernel<<<ks_grid(e, nblk), KS_THREADS, ks_smem_bytes(pg.n_scored), sq>>>(pg, tm, g);
            CU(e, cudaGetLastError());
            e->launches += 1;
        }
        return MB200_OK;
    }
    const bool small_groups = kp.gmax <= KV_GSMALL;
    if (e->fast) {
        if (th == KV_TH_WIDE) {
            if (small_groups) kv_kernel<KV_TH_WIDE, true, KV_GSMALL><<<gv, KV_THREADS, kvb, sq>>>(kp, g);
            else kv_kernel<KV_TH_WIDE, true><<<gv, KV_THREADS, kvb, sq>>>(kp, g);
        } else {
            if (small_groups) kv_kernel<KV_TH_NARROW, true, KV_GSMALL><<<gv, KV_THREADS, kvb, sq>>>(kp, g);
            else kv_kernel<KV_TH_NARROW, true><<<gv, KV_THREADS, kvb, sq>>>(kp, g);
        }
    } else {
        if (th == KV_TH_WIDE) {
            if (small_groups) kv_kernel<KV_TH_WIDE, false, KV_GSMALL><<<gv, KV_THREADS, kvb, sq>>>(kp, g);
            else kv_kernel<KV_TH_WIDE, false><<<gv, KV_THREADS, kvb, sq>>>(kp, g);
        } else {
            if (small_groups) kv_kernel<KV_TH_NARROW, false, KV_GSMALL><<<gv, KV_THREADS, kvb, sq>>>(kp, g);
            else kv_kernel<KV_TH_NARROW, false><<<gv, KV_THREADS, kvb, sq>>>(kp, g);
        }
    }
    CU(e, cudaGetLastError());
    if (after_kv) CU(e, cudaEventRecord(after_kv, sq));
    const int mode = g.dout != nullptr ? KH_DIFF : ((g.dbgG != nullptr || g.dbgL != nullptr) ? KH_DEBUG : KH_MAIN);
    const int ftc = (program == nullptr) ? fused_tc(e) : 0;
    if (mode == KH_MAIN && ftc) {
        // axis-1 pass, DoG and scoring in one kernel: the DoG levels stay in shared memory
        const dim3 gf = kf_grid(e, nblk, ftc);
        if (ftc == 64) {
            if (e->fast) khs_kernel<64, true><<<gf, 256, kf_smem_bytes(64), sq>>>(pg, tm, g);
            else khs_kernel<64, false><<<gf, 256, kf_smem_bytes(64), sq>>>(pg, tm, g);
        } else {
            if (e->fast) khs_kernel<128, true><<<gf, 512, kf_smem_bytes(128), sq>>>(pg, tm, g);
            else khs_kernel<128, false><<<gf, 512, kf_smem_bytes(128), sq>>>(pg, tm, g);
        }
        CU(e, cudaGetLastError());
        if (after_kh) CU(e, cudaEventRecord(after_kh, sq));
        e->launches += 2;
        return MB200_OK;
    }
    // KH_DIFF: difference stack, only the DIFFREF DoGs are kept; KH_DEBUG: dense dumps of mb200_debug_level
    if (e->fast) {
        if (mode == KH_DIFF) kh_kernel<KH_DIFF, true><<<gh, KH_THREADS, khb, sq>>>(pg, tm, g);
        else if (mode == KH_DEBUG) kh_kernel<KH_DEBUG, true><<<gh, KH_THREADS, khb, sq>>>(pg, tm, g);
        else kh_kernel<KH_MAIN, true><<<gh, KH_THREADS, khb, sq>>>(pg, tm, g);
    } else {
        if (mode == KH_DIFF) kh_kernel<KH_DIFF, false><<<gh, KH_THREADS, khb, sq>>>(pg, tm, g);
        else if (mode == KH_DEBUG) kh_kernel<KH_DEBUG, false><<<gh, KH_THREADS, khb, sq>>>(pg, tm, g);
        else kh_kernel<KH_MAIN, false><<<gh, KH_THREADS, khb, sq>>>(pg, tm, g);
    }
    CU(e, cudaGetLastError());
    if (after_kh) CU(e, cudaEventRecord(after_kh, sq));
    e->launches += 2;
    if (pg.n_scored > 0) {
        ks_kernel<<<ks_grid(e, nblk), KS_THREADS, ks_smem_bytes(pg.n_scored), sq>>>(pg, tm, g);
        CU(e, cudaGetLastError());
        e->launches += 1;
    }
    return MB200_OK;
}

int check_block(mb200_engine* e, int block) {
    if (!e) return MB200_ERR_ARG;
    if (!e->configured) return fail(e, MB200_ERR_ARG, "mb200_configure has not been called");
    if (block < 0 || block >= e->nblocks) return fail(e, MB200_ERR_ARG, "block %d out of range [0,%d)", block, e->nblocks);
    return MB200_OK;
}

int refresh_counts(mb200_engine* e) {
    if (e->counts_valid) return MB200_OK;
    if (!e->ran) return fail(e, MB200_ERR_ARG, "mb200_run has not been called for this batch");
    e->h_nz.resize(e->nblocks);
    e->h_rec.resize(e->nblocks);
    e->h_nonfinite.resize(e->nblocks);
    CU(e, cudaMemcpyAsync(e->h_nz.data(), e->nz_count.p, e->nblocks * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaMemcpyAsync(e->h_rec.data(), e->rec_count.p, e->nblocks * sizeof(unsigned long long), cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaMemcpyAsync(e->h_nonfinite.data(), e->nonfinite.p, e->nblocks * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaStreamSynchronize(e->stream));
    e->counts_valid = true;
    return MB200_OK;
}

// Stable segmented LSD radix sort (mb_sort.cuh) of sort_keys[0] / sort_vals[0]; the result is back in buffer 0 (the
// number of passes is rounded up to an even count).  nseg segments of at most max_len keys.
int radix_sort(mb200_engine* e, const RsSegments& sg, int nseg, long long max_len, int key_bits, cudaStream_t sq) {
    if (nseg < 1 || max_len < 1) return MB200_OK;
    const int ntiles = (int)((max_len + RS_TILE - 1) / RS_TILE);
    int st = ensure(e, e->sort_hist, (size_t)nseg * RS_RADIX * (ntiles + 1) * sizeof(unsigned));
    if (st) return st;
    unsigned* hist = (unsigned*)e->sort_hist.p;
    unsigned* tot = hist + (size_t)nseg * RS_RADIX * ntiles;
    int passes = (key_bits + RS_BITS - 1) / RS_BITS;
    passes += passes & 1;
    for (int p = 0; p < passes; ++p) {
        const unsigned long long* kin = (const unsigned long long*)e->sort_keys[p & 1].p;
        unsigned long long* kout = (unsigned long long*)e->sort_keys[(p & 1) ^ 1].p;
        const unsigned* vin = (const unsigned*)e->sort_vals[p & 1].p;
        unsigned* vout = (unsigned*)e->sort_vals[(p & 1) ^ 1].p;
        rs_hist_kernel<<<dim3(ntiles, nseg), RS_THREADS, 0, sq>>>(kin, sg, p * RS_BITS, ntiles, hist);
        rs_scan_tiles_kernel<<<dim3(RS_RADIX, nseg), 256, 0, sq>>>(hist, ntiles, tot);
        rs_scan_digits_kernel<<<nseg, RS_RADIX, 0, sq>>>(tot);
        rs_scatter_kernel<<<dim3(ntiles, nseg), RS_THREADS, 0, sq>>>(kin, vin, kout, vout, sg, p * RS_BITS, ntiles, hist, tot);
        CU(e, cudaGetLastError());
    }
    e->launches += 4 * passes;
    return MB200_OK;
}

}  // namespace

extern "C" {

int mb200_abi_version(void) { return 1; }

int mb200_device_count(int* count) {
    if (!count) return MB200_ERR_ARG;
    cudaError_t st = cudaGetDeviceCount(count);
    if (st != cudaSuccess) {
        *count = 0;
        return MB200_ERR_CUDA;
    }
    return MB200_OK;
}

int mb200_create(int device, mb200_engine** out) {
    if (!out) return MB200_ERR_ARG;
    *out = nullptr;
    int cnt = 0;
    if (cudaGetDeviceCount(&cnt) != cudaSuccess || device < 0 || device >= cnt) return MB200_ERR_CUDA;
    mb200_engine* e = new mb200_engine();
    e->device = device;
    memset(&e->prog, 0, sizeof(e->prog));
    // the upload stream gets the highest priority: its small memset / scatter kernels are placed as soon as an SM has room
    // instead of queueing behind the compute stream's large grids
    int prio_lo = 0, prio_hi = 0;
    if (cudaSetDevice(device) != cudaSuccess || cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi) != cudaSuccess ||
        cudaStreamCreateWithPriority(&e->stream, cudaStreamNonBlocking, prio_lo) != cudaSuccess ||
        cudaStreamCreateWithPriority(&e->up_stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaStreamCreateWithPriority(&e->stream2, cudaStreamNonBlocking, prio_lo) != cudaSuccess ||
        cudaStreamCreateWithPriority(&e->post_stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->ev_post_done, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->ev_chain, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->ev_up, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->ev_run[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->ev_run[1], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreate(&e->ev_begin) != cudaSuccess || cudaEventCreate(&e->ev_prep) != cudaSuccess ||
        cudaEventCreate(&e->ev_end) != cudaSuccess || cudaEventCreate(&e->ev_post0) != cudaSuccess ||
        cudaEventCreate(&e->ev_post1) != cudaSuccess || cudaEventCreate(&e->ev_diff) != cudaSuccess) {
        delete e;
        return MB200_ERR_CUDA;
    }
    *out = e;
    return MB200_OK;
}

void mb200_destroy(mb200_engine* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    if (e->up_stream) cudaStreamSynchronize(e->up_stream);
    DevBuf* all[] = {&e->raw, &e->V, &e->Lb, &e->part_min, &e->part_sum, &e->rec_count, &e->nz_count, &e->nonfinite, &e->rec_row,
                     &e->rec_col, &e->rec_v, &e->rec_sidx, &e->rec_p, &e->fit_loc, &e->fit_scale, &e->st_rows, &e->st_cols,
                     &e->st_vals, &e->st_dense, &e->dbgG, &e->dbgL, &e->rawD, &e->dout, &e->dmu, &e->dsd, &e->rec_pair,
                     &e->d_score_id, &e->d_score_sigma, &e->rec_sid, &e->rec_sigma, &e->d_tmaps, &e->d_dtmaps, &e->nz_xs, &e->nz_ds, &e->nz_perm,
                     &e->nz_vs, &e->nz_out, &e->nz_seg, &e->nz_mean, &e->nz_sd, &e->nz_w, &e->nz_lines,
                     &e->pk_row, &e->pk_col, &e->pk_v, &e->pk_sid, &e->pk_p, &e->pk_sigma, &e->pk_pair, &e->pk_sidx, &e->pk_offsets, &e->st_offsets,
                     &e->nz_x, &e->nz_y, &e->nz_v, &e->sort_keys[0], &e->sort_keys[1], &e->sort_vals[0], &e->sort_vals[1], &e->sort_hist,
                     &e->rec_q, &e->bh_tmin, &e->slotmap, &e->cd_block, &e->cd_row, &e->cd_col, &e->cd_flags, &e->cd_q, &e->cd_sigma,
                     &e->cd_cval, &e->cd_o9, &e->cd_so9, &e->cd_count, &e->cd_slot, &e->cd_pair9, &e->cd_vs9, &e->cd_vo9, &e->dpart, &e->d_pairplan, &e->en_need, &e->en_lines, &e->en_nlist, &e->en_mean,
                     &e->en_scratch};
    for (DevBuf* b : all) release(*b);
    for (cudaEvent_t ev : e->ev_pass) cudaEventDestroy(ev);
    if (e->ev_begin) cudaEventDestroy(e->ev_begin);
    if (e->ev_prep) cudaEventDestroy(e->ev_prep);
    if (e->ev_end) cudaEventDestroy(e->ev_end);
    if (e->ev_post0) cudaEventDestroy(e->ev_post0);
    if (e->ev_post1) cudaEventDestroy(e->ev_post1);
    if (e->ev_diff) cudaEventDestroy(e->ev_diff);
    if (e->ev_up) cudaEventDestroy(e->ev_up);
    for (int k = 0; k < 2; ++k)
        if (e->ev_run[k]) cudaEventDestroy(e->ev_run[k]);
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    if (e->ev_join) cudaEventDestroy(e->ev_join);
    if (e->ev_chain) cudaEventDestroy(e->ev_chain);
    if (e->ev_post_done) cudaEventDestroy(e->ev_post_done);
    if (e->post_stream) { cudaStreamSynchronize(e->post_stream); cudaStreamDestroy(e->post_stream); }
    if (e->stream2) { cudaStreamSynchronize(e->stream2); cudaStreamDestroy(e->stream2); }
    if (e->up_stream) cudaStreamDestroy(e->up_stream);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

const char* mb200_last_error(const mb200_engine* e) { return e ? e->err : "null engine"; }

// Placement of every step's staged box in kh_kernel's shared-memory ring (first fit, wrapping), and for each step the
// latest earlier step whose box it overwrites.
static void plan_ring(const MbProgram& p, int cap, MbStage* stage, int fused_tc = 0) {
    auto size_of = [&](int s) {
        const int w = fused_tc ? kf_box_width(p.st[s].radius, fused_tc) : kh_box_width(p.st[s].radius);
        return (KH_TR * w + 15) & ~15;
    };
    // Boxes small enough for three to fit are packed one after the other, wrapping at the end of the ring.  Larger boxes
    // alternate between the two ENDS of the ring: packed from the bottom, a growing box that wraps to offset 0 reaches into
    // the box of the step before it, and its load can then only start when that step is over (every second step of the
    // largest radii of the 4-octave chain was serialised with its own load that way: 8 % of the axis-1 kernel's samples
    // sat in that one wait).  From the two ends, consecutive boxes overlap only if their sum exceeds the ring.
    // (Each large box takes the end whose latest overlapped box is the older one; ties alternate.)
    auto dep_at = [&](int s, int off, int size) {
        int dep = -1;
        for (int t = 0; t < s; ++t)
            if (stage[t].off < off + size && off < stage[t].off + size_of(t)) dep = t;
        return dep;
    };
    int cur = 0;
    bool top = true;
    for (int s = 0; s < p.n_steps; ++s) {
        const int size = size_of(s);
        int off;
        if (3 * size > cap && size <= cap) {
            const int off_top = (cap - size) & ~15;
            const int dt = dep_at(s, off_top, size), db = dep_at(s, 0, size);
            const bool use_top = dt != db ? dt < db : top;
            off = use_top ? off_top : 0;
            top = !use_top;
            cur = off + size;
        } else {
            if (cur + size > cap) cur = 0;
            off = cur;
            cur += size;
        }
        stage[s].off = off;
        stage[s].dep = dep_at(s, off, size);
    }
}

static void plan_kh_ring(MbProgram& p) {
    plan_ring(p, kh_ring_doubles(p.rmax), p.stage);
    // fused khs_kernel: the 64-column tile (two CTAs per SM) when its ring holds two of the widest boxes, else the
    // 128-column tile (one CTA per SM), else not available (p.pad = tile columns, 0 = none)
    p.pad = kf_fits(p.rmax, 64) ? 64 : (kf_fits(p.rmax, 128) ? 128 : 0);
    if (p.pad) plan_ring(p, kf_ring_doubles(p.pad), p.stage_f, p.pad);
}

// kv_kernel's plan: steps sorted by radius and cut into groups of 1..KV_GMAX consecutive steps.  A group of n steps
// with largest radius R costs R*(2n+1) + n FP64 instructions per output (R pair sums shared by the group, every step
// padded to R taps); the cut minimising the total is found by dynamic programming.
static int plan_kv(mb200_engine* e, const MbProgram& p, KvPlan& kp) {
    memset(&kp, 0, sizeof(kp));
    const int n = p.n_steps;
    std::vector<int> order(n);
    for (int s = 0; s < n; ++s) order[s] = s;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return p.st[a].radius < p.st[b].radius; });
    std::vector<long long> best(n + 1, 0);
    std::vector<int> take(n + 1, 1);
    const int gmax = p.rmax >= KV_GSMALL_RMIN ? KV_GSMALL : KV_GMAX;
    for (int i = n - 1; i >= 0; --i) {
        best[i] = -1;
        for (int c = 1; c <= gmax && i + c <= n; ++c) {
            const long long cost = (long long)p.st[order[i + c - 1]].radius * (2 * c + 1) + c + best[i + c];
            if (best[i] < 0 || cost < best[i]) {
                best[i] = cost;
                take[i] = c;
            }
        }
    }
    kp.rmax = p.rmax;
    kp.gmax = gmax;
    int off = 0;
    for (int pos = 0; pos < n; pos += take[pos]) {
        const int cnt = take[pos];
        if (kp.n_groups >= KV_MAX_GROUPS) return fail(e, MB200_ERR_ARG, "too many steps for the axis-0 plan (%d groups)", KV_MAX_GROUPS);
        KvGroup& gr = kp.grp[kp.n_groups++];
        gr.n = cnt;
        gr.rmax = p.st[order[pos + cnt - 1]].radius;
        gr.tap_off = off;
        const int need = (gr.rmax + 1) * cnt;
        if (off + need > KV_MAX_TAPS_T) return fail(e, MB200_ERR_ARG, "axis-0 plan needs more than %d transposed taps", KV_MAX_TAPS_T);
        for (int slot = 0; slot < cnt; ++slot) {
            const MbStep& st = p.st[order[pos + slot]];
            gr.step[slot] = order[pos + slot];
            for (int j = 0; j <= st.radius; ++j) kp.tapsT[off + j * cnt + slot] = p.taps[st.tap_off + j];
        }
        off += need;
    }
    return MB200_OK;
}

// kvh_kernel's plan: the chain in consecutive pairs (radii never decrease along a chain), each pair sharing its folded pair
// sums over the larger radius, weights transposed and zero-padded as in plan_kv.  kp.n_groups = 0 when the chain does not
// qualify (radius beyond KVH_RMAX, radii not monotone, too many taps).
static void plan_pairs(const MbProgram& p, KvPlan& kp) {
    memset(&kp, 0, sizeof(kp));
    if (p.rmax > KVH_RMAX || (p.n_steps + 1) / 2 > KV_MAX_GROUPS) return;
    int off = 0;
    for (int pos = 0; pos < p.n_steps; pos += 2) {
        const int cnt = std::min(2, p.n_steps - pos);
        KvGroup& gr = kp.grp[kp.n_groups];
        gr.n = cnt;
        gr.rmax = 0;
        for (int slot = 0; slot < cnt; ++slot) gr.rmax = std::max(gr.rmax, p.st[pos + slot].radius);
        gr.tap_off = off;
        const int need = (gr.rmax + 1) * cnt;
        if (off + need > KVH_MAX_TAPS) { kp.n_groups = 0; return; }
        for (int slot = 0; slot < cnt; ++slot) {
            const MbStep& st = p.st[pos + slot];
            gr.step[slot] = pos + slot;
            for (int j = 0; j <= st.radius; ++j) kp.tapsT[off + j * cnt + slot] = p.taps[st.tap_off + j];
        }
        off += need;
        ++kp.n_groups;
    }
    kp.rmax = p.rmax;
}

static int parse_program(mb200_engine* e, MbProgram& p, int n_steps, const int32_t* radius, const int32_t* flags,
                         const int32_t* score_id, const int32_t* tap_off, const double* half_taps, int n_taps) {
    if (!e || !radius || !flags || !score_id || !tap_off || !half_taps) return fail(e, MB200_ERR_ARG, "null argument");
    if (n_steps < 1 || n_steps > MB_MAX_STEPS) return fail(e, MB200_ERR_ARG, "n_steps %d not in [1,%d]", n_steps, MB_MAX_STEPS);
    if (n_taps < 1 || n_taps > MB_MAX_TAPS) return fail(e, MB200_ERR_ARG, "n_taps %d not in [1,%d]", n_taps, MB_MAX_TAPS);
    memset(&p, 0, sizeof(p));
    p.n_steps = n_steps;
    int formed = 0;
    for (int s = 0; s < n_steps; ++s) {
        if (radius[s] < 1 || tap_off[s] < 0 || tap_off[s] + radius[s] + 1 > n_taps)
            return fail(e, MB200_ERR_ARG, "step %d: radius/tap_off out of range", s);
        p.st[s].radius = radius[s];
        p.st[s].tap_off = tap_off[s];
        p.st[s].flags = flags[s];
        p.st[s].score_idx = -1;
        if (s == 0 && !(flags[s] & MB200_STEP_RESTART)) return fail(e, MB200_ERR_ARG, "step 0 must carry MB200_STEP_RESTART");
        formed = (flags[s] & MB200_STEP_RESTART) ? 0 : formed + 1;
        if (flags[s] & MB200_STEP_SCORE) {
            if (formed < 3) return fail(e, MB200_ERR_ARG, "step %d scores before three DoGs of its chain exist", s);
            if (score_id[s] < 1 || score_id[s] > 254) return fail(e, MB200_ERR_ARG, "step %d: score id %d not in [1,254]", s, score_id[s]);
            p.st[s].score_idx = p.n_scored;
            p.score_id[p.n_scored] = score_id[s];
            ++p.n_scored;
        }
        p.rmax = std::max(p.rmax, radius[s]);
    }
    if (p.n_scored > 254) return fail(e, MB200_ERR_ARG, "too many scored steps");
    memcpy(p.taps, half_taps, (size_t)n_taps * sizeof(double));
    plan_kh_ring(p);
    return MB200_OK;
}

int mb200_set_program(mb200_engine* e, int n_steps, const int32_t* radius, const int32_t* flags, const int32_t* score_id,
                      const int32_t* tap_off, const double* half_taps, int n_taps) {
    if (!e) return MB200_ERR_ARG;
    int st = parse_program(e, e->prog, n_steps, radius, flags, score_id, tap_off, half_taps, n_taps);
    if (st) return st;
    if ((st = plan_kv(e, e->prog, e->kvplan))) return st;
    plan_pairs(e->prog, e->pairplan);
    memset(e->score_sigma, 0, sizeof(e->score_sigma));
    e->have_prog = true;
    e->configured = false;
    return MB200_OK;
}

int mb200_set_diff_program(mb200_engine* e, int n_steps, const int32_t* radius, const int32_t* flags, const int32_t* tap_off,
                           const double* half_taps, int n_taps) {
    if (!e) return MB200_ERR_ARG;
    std::vector<int32_t> zero(n_steps > 0 ? n_steps : 1, 0);
    int st = parse_program(e, e->dprog, n_steps, radius, flags, zero.data(), tap_off, half_taps, n_taps);
    if (st) return st;
    int nd = 0;
    for (int s = 0; s < n_steps; ++s) {
        if (flags[s] & MB200_STEP_SCORE) return fail(e, MB200_ERR_ARG, "the difference chain never scores");
        if (flags[s] & MB200_STEP_DIFFREF) {
            if (flags[s] & MB200_STEP_RESTART) return fail(e, MB200_ERR_ARG, "step %d: a MB200_STEP_DIFFREF step must form a DoG", s);
            e->dprog.st[s].score_idx = nd++;         // slot of this step's DoG in dout
        }
    }
    if (nd < 1) return fail(e, MB200_ERR_ARG, "the difference chain needs at least one MB200_STEP_DIFFREF step");
    if ((st = plan_kv(e, e->dprog, e->dkvplan))) return st;
    e->ndiff = nd;
    e->have_dprog = true;
    e->configured = false;          // the TMA descriptors of the difference chain are built by mb200_configure
    return MB200_OK;
}

int mb200_configure(mb200_engine* e, int n, int dpx, int intra, int nblocks, double record_fraction) {
    if (!e) return MB200_ERR_ARG;
    if (!e->have_prog) return fail(e, MB200_ERR_ARG, "mb200_set_program has not been called");
    if (n < 8 || dpx < 1 || nblocks < 1) return fail(e, MB200_ERR_ARG, "bad geometry n=%d dpx=%d nblocks=%d", n, dpx, nblocks);
    if (!intra) return fail(e, MB200_ERR_ARG, "inter-chromosomal tiles are not supported (the reference path is broken, mustache.py:939-942)");
    if (n <= 2 * e->prog.rmax + 1) return fail(e, MB200_ERR_ARG, "tile side %d too small for radius %d", n, e->prog.rmax);
    int st = use_device(e);
    if (st) return st;
    CU(e, cudaStreamWaitEvent(e->stream, e->ev_post_done, 0));    // a post-processing still in flight reads the tile slots
    e->n = n;
    e->dpx = dpx;
    e->intra = intra;
    e->dhi = std::min(dpx + 1, n - 1);
    if (e->dhi < 4) return fail(e, MB200_ERR_ARG, "no diagonal >= 4 in the tile");
    e->wc = e->dhi - 3;
    e->vlo = 2 - e->prog.rmax;
    // odd row lengths: the skewed TMA view has row stride (len-1)*8 bytes, which must be a multiple of 16
    e->wv = (e->dhi + 2 * e->prog.rmax + 1) | 1;
    e->nblocks = nblocks;
    const double frac = record_fraction > 0 ? record_fraction : 0.125;
    e->rec_cap = std::max<long long>(4096, (long long)(frac * (double)n * e->wc));
    e->wl = (e->dhi + 1) | 1;                            // diagonals 2..dhi+2, odd row length (skewed TMA view)
    e->plane_v = ((long long)n * e->wv + 1) & ~1LL;      // plane strides: multiples of 16 bytes
    e->plane_l = ((long long)n * e->wl + 1) & ~1LL;
    dim3 gh = ks_grid(e, 1);
    e->ncta_h = gh.x * gh.y;
    if (fused_tc(e)) {                                   // the fused kernel writes one partial per warp
        dim3 gf = kf_grid(e, 1, fused_tc(e));
        e->ncta_h = gf.x * gf.y * (fused_tc(e) / KS_K);
    }
    if ((st = set_smem_limits(e))) return st;
    const size_t ns = (size_t)std::max(e->prog.n_scored, 1);
    const size_t B = nblocks;
    CU(e, cudaStreamSynchronize(e->stream));             // a new geometry re-uses every buffer
    CU(e, cudaStreamSynchronize(e->up_stream));
    if ((st = ensure(e, e->raw, 2 * B * n * e->wc * sizeof(double)))) return st;
    e->slot_up = e->slot_run = 0;
    e->slot_used[0] = e->slot_used[1] = false;
    e->up_dirty = false;
    if ((st = ensure(e, e->part_min, B * ns * e->ncta_h * sizeof(double)))) return st;
    if ((st = ensure(e, e->part_sum, B * ns * e->ncta_h * sizeof(double)))) return st;
    if ((st = ensure(e, e->rec_count, B * sizeof(unsigned long long)))) return st;
    if ((st = ensure(e, e->nz_count, B * sizeof(unsigned long long)))) return st;
    if ((st = ensure(e, e->nonfinite, B * sizeof(int)))) return st;
    if ((st = ensure(e, e->rec_row, B * e->rec_cap * sizeof(int)))) return st;
    if ((st = ensure(e, e->rec_col, B * e->rec_cap * sizeof(int)))) return st;
    if ((st = ensure(e, e->rec_sidx, B * e->rec_cap * sizeof(int)))) return st;
    if ((st = ensure(e, e->rec_v, B * e->rec_cap * sizeof(double)))) return st;
    if ((st = ensure(e, e->rec_p, B * e->rec_cap * sizeof(double)))) return st;
    if ((st = ensure(e, e->rec_sid, B * e->rec_cap * sizeof(int)))) return st;
    if ((st = ensure(e, e->rec_sigma, B * e->rec_cap * sizeof(double)))) return st;
    if ((st = ensure(e, e->d_score_id, MB_MAX_STEPS * sizeof(int)))) return st;
    if ((st = ensure(e, e->d_score_sigma, MB_MAX_STEPS * sizeof(double)))) return st;
    if ((st = ensure(e, e->fit_loc, B * ns * sizeof(double)))) return st;
    if ((st = ensure(e, e->fit_scale, B * ns * sizeof(double)))) return st;
    if (e->have_dprog && nblocks % 2 == 0 && e->dprog.rmax <= e->prog.rmax) {
        // differential batches (blocks 2k / 2k+1 = the two maps of pair k): difference tile, its kept DoGs, norm.fit and
        // pPair per record -- allocated here so that the scratch below is sized from what is really left
        const size_t tile = (size_t)n * e->wc, np = B / 2;
        if ((st = ensure(e, e->rawD, np * tile * sizeof(double)))) return st;
        if ((st = ensure(e, e->dout, (size_t)e->ndiff * np * tile * sizeof(double)))) return st;
        if ((st = ensure(e, e->dmu, (size_t)e->ndiff * np * sizeof(double)))) return st;
        if ((st = ensure(e, e->dsd, (size_t)e->ndiff * np * sizeof(double)))) return st;
        if ((st = ensure(e, e->rec_pair, B * e->rec_cap * sizeof(double)))) return st;
    }
    // axis-0 scratch: as many blocks per pass as fit in ~80 % of what is free now (plus what V already holds)
    size_t free_b = 0, total_b = 0;
    CU(e, cudaMemGetInfo(&free_b, &total_b));
    const size_t per_block = v_bytes_per_block(e) + l_bytes_per_block(e);
    const size_t budget = (size_t)(0.8 * (double)(free_b + e->V.cap + e->Lb.cap)) - 4 * V_GUARD_BYTES;
    long long fit = (long long)(budget / per_block);
    if (fit < 1) return fail(e, MB200_ERR_NOMEM, "axis-0 scratch for one block needs %zu bytes, %zu available", per_block, budget);
    e->pass_blocks = (int)std::min<long long>(fit, nblocks);
    if (e->pass_limit > 0) e->pass_blocks = std::min(e->pass_blocks, e->pass_limit);
    if ((st = ensure(e, e->V, (size_t)e->pass_blocks * v_bytes_per_block(e) + 2 * V_GUARD_BYTES))) return st;
    if ((st = ensure(e, e->Lb, (size_t)e->pass_blocks * l_bytes_per_block(e) + 2 * V_GUARD_BYTES))) return st;
    CU(e, cudaMemsetAsync(e->raw.p, 0, 2 * B * n * e->wc * sizeof(double), e->stream));
    if ((st = encode_maps(e, e->prog, e->tmaps, true))) return st;
    if ((st = ensure(e, e->d_tmaps, sizeof(MbTensorMaps)))) return st;
    CU(e, cudaMemcpyAsync(e->d_tmaps.p, &e->tmaps, sizeof(MbTensorMaps), cudaMemcpyHostToDevice, e->stream));
    if (e->have_dprog && e->dprog.rmax <= e->prog.rmax) {
        if ((st = encode_maps(e, e->dprog, e->dtmaps, false))) return st;
        if ((st = ensure(e, e->d_dtmaps, sizeof(MbTensorMaps)))) return st;
        CU(e, cudaMemcpyAsync(e->d_dtmaps.p, &e->dtmaps, sizeof(MbTensorMaps), cudaMemcpyHostToDevice, e->stream));
    }
    if ((st = ensure(e, e->d_pairplan, sizeof(KvPlan)))) return st;
    CU(e, cudaMemcpyAsync(e->d_pairplan.p, &e->pairplan, sizeof(KvPlan), cudaMemcpyHostToDevice, e->stream));
    CU(e, cudaStreamSynchronize(e->stream));      // the host copies may be re-encoded by the next configure; tiles are zeroed
    e->configured = true;
    e->ran = false;
    e->counts_valid = false;
    e->packed = false;
    e->post_done = false;
    return MB200_OK;
}

int mb200_upload_coo_host(mb200_engine* e, int block, const int32_t* rows, const int32_t* cols, const double* vals,
                          int64_t nnz) {
    int st = check_block(e, block);
    if (st) return st;
    if (nnz < 0 || (nnz > 0 && (!rows || !cols || !vals))) return fail(e, MB200_ERR_ARG, "bad COO arguments");
    if ((st = use_device(e))) return st;
    if (nnz == 0) {
        if ((st = begin_upload(e))) return st;
        CU(e, cudaMemsetAsync(raw_slot(e, e->slot_up) + (size_t)block * e->n * e->wc, 0, (size_t)e->n * e->wc * sizeof(double), e->up_stream));
        return MB200_OK;
    }
    if ((st = ensure(e, e->st_rows, nnz * sizeof(int)))) return st;
    if ((st = ensure(e, e->st_cols, nnz * sizeof(int)))) return st;
    if ((st = ensure(e, e->st_vals, nnz * sizeof(double)))) return st;
    if ((st = begin_upload(e))) return st;
    CU(e, cudaMemcpyAsync(e->st_rows.p, rows, nnz * sizeof(int), cudaMemcpyHostToDevice, e->up_stream));
    CU(e, cudaMemcpyAsync(e->st_cols.p, cols, nnz * sizeof(int), cudaMemcpyHostToDevice, e->up_stream));
    CU(e, cudaMemcpyAsync(e->st_vals.p, vals, nnz * sizeof(double), cudaMemcpyHostToDevice, e->up_stream));
    double* rawb = raw_slot(e, e->slot_up) + (size_t)block * e->n * e->wc;
    CU(e, cudaMemsetAsync(rawb, 0, (size_t)e->n * e->wc * sizeof(double), e->up_stream));   // the slot may hold an older tile
    const int grid = (int)std::min<int64_t>((nnz + 255) / 256, 148 * 8);
    scatter_coo_kernel<<<grid, 256, 0, e->up_stream>>>((const int*)e->st_rows.p, (const int*)e->st_cols.p,
                                                       (const double*)e->st_vals.p, nnz, rawb, e->n, e->wc, e->dhi);
    CU(e, cudaGetLastError());
    // The staging buffers are reused by the next upload; uploads are ordered on up_stream, so the scatter above has read
    // them before the next copies land.  The host arrays are pageable (numpy): cudaMemcpyAsync returns once they have been
    // staged, so they may be released when this call returns; no stream synchronisation per block.
    return MB200_OK;
}

int mb200_upload_coo_batch(mb200_engine* e, int first_block, int nblk, const int64_t* offsets, const int32_t* rows,
                           const int32_t* cols, const double* vals) {
    int st = check_block(e, first_block);
    if (st) return st;
    if (nblk < 1 || first_block + nblk > e->nblocks || !offsets) return fail(e, MB200_ERR_ARG, "bad block range");
    const int64_t nnz = offsets[nblk];
    if (offsets[0] != 0 || nnz < 0 || (nnz > 0 && (!rows || !cols || !vals))) return fail(e, MB200_ERR_ARG, "bad COO arguments");
    for (int b = 0; b < nblk; ++b)
        if (offsets[b + 1] < offsets[b]) return fail(e, MB200_ERR_ARG, "offsets must not decrease");
    if ((st = use_device(e))) return st;
    if ((st = ensure(e, e->st_rows, std::max<int64_t>(nnz, 1) * sizeof(int)))) return st;
    if ((st = ensure(e, e->st_cols, std::max<int64_t>(nnz, 1) * sizeof(int)))) return st;
    if ((st = ensure(e, e->st_vals, std::max<int64_t>(nnz, 1) * sizeof(double)))) return st;
    if ((st = ensure(e, e->st_offsets, (size_t)(e->nblocks + 1) * sizeof(long long)))) return st;
    if ((st = begin_upload(e))) return st;
    double* raw0 = raw_slot(e, e->slot_up) + (size_t)first_block * e->n * e->wc;
    CU(e, cudaMemsetAsync(raw0, 0, (size_t)nblk * e->n * e->wc * sizeof(double), e->up_stream));
    if (nnz == 0) return MB200_OK;
    e->h_offsets.assign(offsets, offsets + nblk + 1);                 // engine-owned copy: the caller's may go away
    CU(e, cudaMemcpyAsync(e->st_offsets.p, e->h_offsets.data(), (size_t)(nblk + 1) * sizeof(long long), cudaMemcpyHostToDevice, e->up_stream));
    CU(e, cudaMemcpyAsync(e->st_rows.p, rows, nnz * sizeof(int), cudaMemcpyHostToDevice, e->up_stream));
    CU(e, cudaMemcpyAsync(e->st_cols.p, cols, nnz * sizeof(int), cudaMemcpyHostToDevice, e->up_stream));
    CU(e, cudaMemcpyAsync(e->st_vals.p, vals, nnz * sizeof(double), cudaMemcpyHostToDevice, e->up_stream));
    const int grid = (int)std::min<int64_t>((nnz + 255) / 256, 148 * 8);
    scatter_coo_batch_kernel<<<grid, 256, 0, e->up_stream>>>((const int*)e->st_rows.p, (const int*)e->st_cols.p,
                                                             (const double*)e->st_vals.p, (const long long*)e->st_offsets.p, nblk,
                                                             raw0, e->n, e->wc, e->dhi);
    CU(e, cudaGetLastError());
    return MB200_OK;
}

int mb200_upload_coo_dev(mb200_engine* e, int block, const int32_t* rows_dev, const int32_t* cols_dev, const double* vals_dev,
                         int64_t nnz) {
    int st = check_block(e, block);
    if (st) return st;
    if (nnz < 0 || (nnz > 0 && (!rows_dev || !cols_dev || !vals_dev))) return fail(e, MB200_ERR_ARG, "bad COO arguments");
    if ((st = use_device(e))) return st;
    if ((st = begin_upload(e))) return st;
    double* rawb = raw_slot(e, e->slot_up) + (size_t)block * e->n * e->wc;
    CU(e, cudaMemsetAsync(rawb, 0, (size_t)e->n * e->wc * sizeof(double), e->up_stream));
    if (nnz == 0) return MB200_OK;
    const int grid = (int)std::min<int64_t>((nnz + 255) / 256, 148 * 8);
    scatter_coo_kernel<<<grid, 256, 0, e->up_stream>>>(rows_dev, cols_dev, vals_dev, nnz, rawb, e->n, e->wc, e->dhi);
    CU(e, cudaGetLastError());
    return MB200_OK;
}

int mb200_upload_band_host(mb200_engine* e, int block, const double* band, int64_t wsrc) {
    int st = check_block(e, block);
    if (st) return st;
    if (!band || wsrc < 1) return fail(e, MB200_ERR_ARG, "bad band arguments");
    if ((st = use_device(e))) return st;
    if ((st = begin_upload(e))) return st;
    double* rawb = raw_slot(e, e->slot_up) + (size_t)block * e->n * e->wc;
    const size_t w = (size_t)std::min<int64_t>(wsrc, e->wc);
    if (w < (size_t)e->wc) CU(e, cudaMemsetAsync(rawb, 0, (size_t)e->n * e->wc * sizeof(double), e->up_stream));
    CU(e, cudaMemcpy2DAsync(rawb, (size_t)e->wc * sizeof(double), band, (size_t)wsrc * sizeof(double), w * sizeof(double),
                            e->n, cudaMemcpyHostToDevice, e->up_stream));
    return MB200_OK;
}

int mb200_upload_dense_host(mb200_engine* e, int block, const double* tile, int64_t ld) {
    int st = check_block(e, block);
    if (st) return st;
    if (!tile || ld < e->n) return fail(e, MB200_ERR_ARG, "bad dense arguments");
    if ((st = use_device(e))) return st;
    if ((st = begin_upload(e))) return st;
    double* rawb = raw_slot(e, e->slot_up) + (size_t)block * e->n * e->wc;
    // Row i of the band starts at tile[i*ld + i + 4]: a pitched copy with source pitch (ld+1) moves exactly the band.
    // Rows whose band segment would run past the end of the host array are copied one by one, clipped.
    // The band tails of the clipped rows below (columns >= n) are never written by ANY upload path (the scatter kernels
    // check c < n), so they still hold the zeros mb200_configure put there; every other band element is overwritten here.
    // No memset: a memset kernel queued behind the compute stream's grids would delay the copies that follow it.
    const int64_t total = (int64_t)(e->n - 1) * ld + e->n;                 // elements addressable in the host tile
    int64_t safe_rows = (total - 4 - e->wc) / (ld + 1) + 1;               // rows i with i*(ld+1) + 4 + wc <= total
    safe_rows = std::max<int64_t>(0, std::min<int64_t>(safe_rows, e->n));
    if (safe_rows > 0)
        CU(e, cudaMemcpy2DAsync(rawb, (size_t)e->wc * sizeof(double), tile + 4, (size_t)(ld + 1) * sizeof(double),
                                (size_t)e->wc * sizeof(double), (size_t)safe_rows, cudaMemcpyHostToDevice, e->up_stream));
    for (int64_t i = safe_rows; i < e->n; ++i) {
        const int64_t w = std::min<int64_t>(e->wc, e->n - i - 4);
        if (w > 0)
            CU(e, cudaMemcpyAsync(rawb + (size_t)i * e->wc, tile + i * ld + i + 4, (size_t)w * sizeof(double),
                                  cudaMemcpyHostToDevice, e->up_stream));
    }
    return MB200_OK;
}

int mb200_upload_dense_dev(mb200_engine* e, int block, const double* tile_dev, int64_t ld) {
    int st = check_block(e, block);
    if (st) return st;
    if (!tile_dev || ld < e->n) return fail(e, MB200_ERR_ARG, "bad dense arguments");
    if ((st = use_device(e))) return st;
    if ((st = begin_upload(e))) return st;
    double* rawb = raw_slot(e, e->slot_up) + (size_t)block * e->n * e->wc;
    band_from_dense_kernel<<<148 * 8, 256, 0, e->up_stream>>>(tile_dev, ld, rawb, e->n, e->wc);
    CU(e, cudaGetLastError());
    return MB200_OK;
}

int mb200_run(mb200_engine* e) {
    if (!e) return MB200_ERR_ARG;
    if (!e->configured) return fail(e, MB200_ERR_ARG, "mb200_configure has not been called");
    int st = use_device(e);
    if (st) return st;
    const int B = e->nblocks;
    e->launches = 0;
    e->counts_valid = false;
    e->ran_diff = false;
    e->packed = false;
    e->post_done = false;
    CU(e, cudaStreamWaitEvent(e->stream, e->ev_post_done, 0));    // the post-processing of the previous batch reads its records
    if ((st = adopt_uploads(e))) return st;
    // Passes: as many blocks as the scratch holds.  Overlapped mode: the scratch is split into two regions and the passes
    // alternate between two streams, so that the (issue-bound) scoring of one half of the batch runs while the (FP64 /
    // HBM-bound) Gaussian passes of the other half do.
    const bool two = e->overlap && B >= 2 && e->pass_blocks >= 2;
    // overlapped: up to 8 passes, pipelined -- pass p+1 starts its axis-0 kernel when pass p has finished its axis-1
    // kernel, so the scoring of pass p runs next to the Gaussian passes of pass p+1
    const int per_pass = two ? std::max(1, std::min(e->pass_blocks / 2, (B + 7) / 8)) : e->pass_blocks;
    const int npass = (B + per_pass - 1) / per_pass;
    e->npass_run = npass;
    while ((int)e->ev_pass.size() < 4 * npass) {
        cudaEvent_t ev;
        CU(e, cudaEventCreate(&ev));
        e->ev_pass.push_back(ev);
    }
    CU(e, cudaEventRecord(e->ev_begin, e->stream));
    CU(e, cudaMemsetAsync(e->rec_count.p, 0, B * sizeof(unsigned long long), e->stream));
    CU(e, cudaMemsetAsync(e->nz_count.p, 0, B * sizeof(unsigned long long), e->stream));
    CU(e, cudaMemsetAsync(e->nonfinite.p, 0, B * sizeof(int), e->stream));
    count_mask_kernel<<<dim3(B >= 8 ? 148 * 2 : 148 * 8, B), 256, 0, e->stream>>>(raw_slot(e, e->slot_run), e->n, e->wc,
                                                                (unsigned long long*)e->nz_count.p, (int*)e->nonfinite.p);
    CU(e, cudaGetLastError());
    e->launches += 1;
    CU(e, cudaEventRecord(e->ev_prep, e->stream));
    if (two) {
        CU(e, cudaEventRecord(e->ev_fork, e->stream));
        CU(e, cudaStreamWaitEvent(e->stream2, e->ev_fork, 0));
    }
    for (int p = 0; p < npass; ++p) {
        const int first = p * per_pass, nb = std::min(per_pass, B - first);
        cudaStream_t sq = (two && (p & 1)) ? e->stream2 : e->stream;
        const int zoff = (two && (p & 1)) ? e->pass_blocks / 2 : 0;
        if (two && p > 0) CU(e, cudaStreamWaitEvent(sq, e->ev_pass[4 * (p - 1) + 2], 0));
        CU(e, cudaEventRecord(e->ev_pass[4 * p], sq));
        if ((st = launch_pass(e, first, nb, nullptr, e->ev_pass[4 * p + 1], nullptr, e->ev_pass[4 * p + 2], sq, zoff))) return st;
        CU(e, cudaEventRecord(e->ev_pass[4 * p + 3], sq));
    }
    if (two) {
        CU(e, cudaEventRecord(e->ev_join, e->stream2));
        CU(e, cudaStreamWaitEvent(e->stream, e->ev_join, 0));
    }
    if (e->prog.n_scored > 0) {
        reduce_stats_kernel<<<dim3(e->prog.n_scored, B), 256, 0, e->stream>>>(
            (const double*)e->part_min.p, (const double*)e->part_sum.p, e->ncta_h, e->prog.n_scored,
            (const unsigned long long*)e->nz_count.p, (double*)e->fit_loc.p, (double*)e->fit_scale.p);
        CU(e, cudaGetLastError());
        // engine-owned host tables: safe to copy asynchronously (they only change in set_program / set_score_sigmas,
        // which are not legal while a run is in flight)
        CU(e, cudaMemcpyAsync(e->d_score_id.p, e->prog.score_id, MB_MAX_STEPS * sizeof(int), cudaMemcpyHostToDevice, e->stream));
        CU(e, cudaMemcpyAsync(e->d_score_sigma.p, e->score_sigma, MB_MAX_STEPS * sizeof(double), cudaMemcpyHostToDevice, e->stream));
        finalise_kernel<<<dim3(B >= 8 ? 64 : 148 * 4, B), 256, 0, e->stream>>>(
            (const unsigned long long*)e->rec_count.p, e->rec_cap, (const double*)e->rec_v.p, (const int*)e->rec_sidx.p,
            e->prog.n_scored, (const double*)e->fit_loc.p, (const double*)e->fit_scale.p, (const int*)e->d_score_id.p,
            (const double*)e->d_score_sigma.p, (double*)e->rec_p.p, (int*)e->rec_sid.p, (double*)e->rec_sigma.p);
        CU(e, cudaGetLastError());
        e->launches += 2;
    }
    CU(e, cudaEventRecord(e->ev_end, e->stream));
    CU(e, cudaEventRecord(e->ev_run[e->slot_run], e->stream));
    e->slot_used[e->slot_run] = true;
    e->ran = true;
    return MB200_OK;
}

int mb200_run_after(mb200_engine* e, mb200_engine* other) {
    if (!e || !other || e == other) return MB200_ERR_ARG;
    if (e->device != other->device) return fail(e, MB200_ERR_ARG, "mb200_run_after needs both engines on one device");
    int st = use_device(e);
    if (st) return st;
    CU(e, cudaEventRecord(other->ev_chain, other->stream));
    CU(e, cudaStreamWaitEvent(e->stream, other->ev_chain, 0));
    return MB200_OK;
}

int mb200_sync(mb200_engine* e) {
    if (!e) return MB200_ERR_ARG;
    int st = use_device(e);
    if (st) return st;
    CU(e, cudaStreamSynchronize(e->up_stream));
    CU(e, cudaStreamSynchronize(e->stream2));
    CU(e, cudaStreamSynchronize(e->stream));
    CU(e, cudaStreamSynchronize(e->post_stream));
    return MB200_OK;
}

int mb200_block_counts(mb200_engine* e, int block, int64_t* nz_count, int64_t* n_found) {
    int st = check_block(e, block);
    if (st) return st;
    if ((st = use_device(e))) return st;
    if ((st = refresh_counts(e))) return st;
    if (nz_count) *nz_count = (int64_t)e->h_nz[block];
    if (n_found) *n_found = (int64_t)e->h_rec[block];
    if (e->h_nonfinite[block]) return fail(e, MB200_ERR_NONFINITE, "block %d holds non-finite values", block);
    if ((long long)e->h_rec[block] > e->rec_cap)
        return fail(e, MB200_ERR_CAPACITY, "block %d produced %llu records, capacity %lld", block, e->h_rec[block], e->rec_cap);
    return MB200_OK;
}

int mb200_fetch_records(mb200_engine* e, int block, int64_t capacity, int32_t* rows, int32_t* cols, double* v,
                        int32_t* score_id, double* p, int64_t* n_out) {
    int64_t nz = 0, nf = 0;
    int st = mb200_block_counts(e, block, &nz, &nf);
    if (n_out) *n_out = nf;
    if (st) return st;
    const int64_t m = std::min<int64_t>(nf, capacity);
    if (m <= 0) return MB200_OK;
    if (!rows || !cols || !v || !score_id || !p) return fail(e, MB200_ERR_ARG, "null output array");
    const size_t o = (size_t)block * e->rec_cap;
    CU(e, cudaMemcpyAsync(rows, (int*)e->rec_row.p + o, m * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaMemcpyAsync(cols, (int*)e->rec_col.p + o, m * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaMemcpyAsync(score_id, (int*)e->rec_sid.p + o, m * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaMemcpyAsync(v, (double*)e->rec_v.p + o, m * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaMemcpyAsync(p, (double*)e->rec_p.p + o, m * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaStreamSynchronize(e->stream));
    return MB200_OK;
}

int mb200_set_arithmetic(mb200_engine* e, int fused_multiply_add) {
    if (!e || fused_multiply_add < 0 || fused_multiply_add > 1) return MB200_ERR_ARG;
    e->fast = fused_multiply_add;
    return MB200_OK;
}

int mb200_set_fusion(mb200_engine* e, int enable) {
    if (!e || enable < 0 || enable > 2) return MB200_ERR_ARG;
    if (e->fusion != enable) e->configured = false;     // the partial-statistics layout depends on the path
    e->fusion = enable;
    return MB200_OK;
}

int mb200_set_overlap(mb200_engine* e, int enable) {
    if (!e || enable < 0 || enable > 1) return MB200_ERR_ARG;
    e->overlap = enable;
    return MB200_OK;
}

int mb200_set_pass_limit(mb200_engine* e, int max_blocks) {
    if (!e || max_blocks < 0) return MB200_ERR_ARG;
    e->pass_limit = max_blocks;
    e->configured = false;
    return MB200_OK;
}

// Packs the records of every block of the batch into contiguous arrays (block b at [offsets[b], offsets[b+1])), on the
// device: one kernel, so that a batch is fetched (or handed to NCCL) with one copy per field instead of one per block.
int mb200_pack_records(mb200_engine* e, int64_t* offsets, int64_t* total) {
    if (!e) return MB200_ERR_ARG;
    if (!e->configured) return fail(e, MB200_ERR_ARG, "mb200_configure has not been called");
    int st = use_device(e);
    if (st) return st;
    if ((st = refresh_counts(e))) return st;
    const int B = e->nblocks;
    e->pk_off.assign(B + 1, 0);
    for (int b = 0; b < B; ++b) {
        if (e->h_nonfinite[b]) return fail(e, MB200_ERR_NONFINITE, "block %d holds non-finite values", b);
        if ((long long)e->h_rec[b] > e->rec_cap)
            return fail(e, MB200_ERR_CAPACITY, "block %d produced %llu records, capacity %lld", b, e->h_rec[b], e->rec_cap);
        e->pk_off[b + 1] = e->pk_off[b] + (long long)e->h_rec[b];
    }
    const long long tot = e->pk_off[B];
    if (offsets)
        for (int b = 0; b <= B; ++b) offsets[b] = e->pk_off[b];
    if (total) *total = tot;
    if (!e->packed && tot > 0) {
        const size_t m = (size_t)tot;
        if ((st = ensure(e, e->pk_row, m * sizeof(int)))) return st;
        if ((st = ensure(e, e->pk_col, m * sizeof(int)))) return st;
        if ((st = ensure(e, e->pk_sid, m * sizeof(int)))) return st;
        if ((st = ensure(e, e->pk_sidx, m * sizeof(int)))) return st;
        if ((st = ensure(e, e->pk_v, m * sizeof(double)))) return st;
        if ((st = ensure(e, e->pk_p, m * sizeof(double)))) return st;
        if ((st = ensure(e, e->pk_sigma, m * sizeof(double)))) return st;
        if (e->ran_diff && (st = ensure(e, e->pk_pair, m * sizeof(double)))) return st;
        if ((st = ensure(e, e->pk_offsets, (size_t)(B + 1) * sizeof(long long)))) return st;
        CU(e, cudaMemcpyAsync(e->pk_offsets.p, e->pk_off.data(), (size_t)(B + 1) * sizeof(long long), cudaMemcpyHostToDevice, e->stream));
        const long long avg = tot / B + 1;
        const int gx = (int)std::max<long long>(1, std::min<long long>((avg + 255) / 256, B >= 8 ? 32 : 148 * 4));
        pack_records_kernel<<<dim3(gx, B), 256, 0, e->stream>>>(
            (const long long*)e->pk_offsets.p, e->rec_cap, (const int*)e->rec_row.p, (const int*)e->rec_col.p, (const double*)e->rec_v.p,
            (const int*)e->rec_sid.p, (const int*)e->rec_sidx.p, (const double*)e->rec_p.p, (const double*)e->rec_sigma.p,
            e->ran_diff ? (const double*)e->rec_pair.p : nullptr, (int*)e->pk_row.p, (int*)e->pk_col.p, (double*)e->pk_v.p,
            (int*)e->pk_sid.p, (int*)e->pk_sidx.p, (double*)e->pk_p.p, (double*)e->pk_sigma.p, (double*)e->pk_pair.p);
        CU(e, cudaGetLastError());
        e->launches += 1;
    }
    e->packed = true;
    return MB200_OK;
}

int mb200_packed_device(mb200_engine* e, void** rows, void** cols, void** v, void** score_id, void** scored_index, void** p,
                        void** sigma, void** pair) {
    if (!e) return MB200_ERR_ARG;
    if (!e->packed) return fail(e, MB200_ERR_ARG, "mb200_pack_records has not been called for this batch");
    if (rows) *rows = e->pk_row.p;
    if (cols) *cols = e->pk_col.p;
    if (v) *v = e->pk_v.p;
    if (score_id) *score_id = e->pk_sid.p;
    if (scored_index) *scored_index = e->pk_sidx.p;
    if (p) *p = e->pk_p.p;
    if (sigma) *sigma = e->pk_sigma.p;
    if (pair) *pair = e->ran_diff ? e->pk_pair.p : nullptr;
    return MB200_OK;
}

int mb200_fetch_packed(mb200_engine* e, int64_t capacity, int32_t* rows, int32_t* cols, double* v, int32_t* score_id, double* p,
                       double* sigma, double* pair) {
    if (!e) return MB200_ERR_ARG;
    if (!e->packed) return fail(e, MB200_ERR_ARG, "mb200_pack_records has not been called for this batch");
    int st = use_device(e);
    if (st) return st;
    const long long tot = e->pk_off.back();
    if (tot > capacity) return fail(e, MB200_ERR_CAPACITY, "%lld records, caller capacity %lld", tot, (long long)capacity);
    if (pair && !e->ran_diff) return fail(e, MB200_ERR_ARG, "mb200_run_differential has not been called for this batch");
    if (tot > 0) {
        const size_t m = (size_t)tot;
        if (rows) CU(e, cudaMemcpyAsync(rows, e->pk_row.p, m * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
        if (cols) CU(e, cudaMemcpyAsync(cols, e->pk_col.p, m * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
        if (score_id) CU(e, cudaMemcpyAsync(score_id, e->pk_sid.p, m * sizeof(int), cudaMemcpyDeviceToHost, e->stream));
        if (v) CU(e, cudaMemcpyAsync(v, e->pk_v.p, m * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
        if (p) CU(e, cudaMemcpyAsync(p, e->pk_p.p, m * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
        if (sigma) CU(e, cudaMemcpyAsync(sigma, e->pk_sigma.p, m * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
        if (pair) CU(e, cudaMemcpyAsync(pair, e->pk_pair.p, m * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    }
    CU(e, cudaStreamSynchronize(e->stream));
    return MB200_OK;
}

int mb200_batch_counts(mb200_engine* e, int64_t* nz_count, int64_t* n_found) {
    if (!e) return MB200_ERR_ARG;
    if (!e->configured) return fail(e, MB200_ERR_ARG, "mb200_configure has not been called");
    int st = use_device(e);
    if (st) return st;
    if ((st = refresh_counts(e))) return st;
    for (int b = 0; b < e->nblocks; ++b) {
        if (nz_count) nz_count[b] = (int64_t)e->h_nz[b];
        if (n_found) n_found[b] = (int64_t)e->h_rec[b];
    }
    return MB200_OK;
}

// Device half of the block post-processing (mb_post.cuh): BH per block, o < pt, sparsity filter, neighbourhood patches.
int mb200_select_candidates(mb200_engine* e, double pt, double st_thr, double candidate_fraction) {
    if (!e) return MB200_ERR_ARG;
    if (!e->configured || !e->ran) return fail(e, MB200_ERR_ARG, "mb200_run has not been called for this batch");
    if (!(pt <= 1.0)) return fail(e, MB200_ERR_ARG, "pt > 1 would select off-mask pixels in the reference; not supported");
    int st = use_device(e);
    if (st) return st;
    const int B = e->nblocks;
    const size_t slots = (size_t)B * e->rec_cap;
    cudaStream_t sq = e->post_stream;                 // the run is complete once refresh_counts below has returned
    // one small device -> host read of the record counters: the sort and BH grids are sized by the fullest block, not by
    // the capacity, and an overflow or a non-finite tile is reported before any work is queued
    if ((st = refresh_counts(e))) return st;
    long long max_found = 1, total_found = 0;
    for (int b = 0; b < B; ++b) {
        if (e->h_nonfinite[b]) return fail(e, MB200_ERR_NONFINITE, "block %d holds non-finite values", b);
        if ((long long)e->h_rec[b] > e->rec_cap)
            return fail(e, MB200_ERR_CAPACITY, "block %d produced %llu records, capacity %lld", b, e->h_rec[b], e->rec_cap);
        max_found = std::max<long long>(max_found, (long long)e->h_rec[b]);
        total_found += (long long)e->h_rec[b];
    }
    for (int k = 0; k < 2; ++k) {
        if ((st = ensure(e, e->sort_keys[k], slots * sizeof(unsigned long long)))) return st;
        if ((st = ensure(e, e->sort_vals[k], slots * sizeof(unsigned)))) return st;
    }
    if ((st = ensure(e, e->rec_q, slots * sizeof(double)))) return st;
    const int ntiles = (int)((max_found + RS_TILE - 1) / RS_TILE);
    if ((st = ensure(e, e->bh_tmin, (size_t)B * ntiles * sizeof(double)))) return st;
    const size_t tile = (size_t)e->n * e->wc;
    if ((st = ensure(e, e->slotmap, (size_t)B * tile * sizeof(int)))) return st;
    const double frac = candidate_fraction > 0 ? candidate_fraction : 1.0 / 16.0;
    e->cand_cap = std::max<long long>(16, (long long)(frac * (double)total_found) + 1);
    const size_t cc = (size_t)e->cand_cap;
    if ((st = ensure(e, e->cd_block, cc * sizeof(int)))) return st;
    if ((st = ensure(e, e->cd_row, cc * sizeof(int)))) return st;
    if ((st = ensure(e, e->cd_col, cc * sizeof(int)))) return st;
    if ((st = ensure(e, e->cd_flags, cc * sizeof(int)))) return st;
    if ((st = ensure(e, e->cd_q, cc * sizeof(double)))) return st;
    if ((st = ensure(e, e->cd_sigma, cc * sizeof(double)))) return st;
    if ((st = ensure(e, e->cd_cval, cc * sizeof(double)))) return st;
    if ((st = ensure(e, e->cd_o9, cc * 9 * sizeof(double)))) return st;
    if ((st = ensure(e, e->cd_so9, cc * 9 * sizeof(double)))) return st;
    if ((st = ensure(e, e->cd_count, sizeof(unsigned long long)))) return st;
    if ((st = ensure(e, e->cd_slot, cc * sizeof(int)))) return st;
    e->post_diff = e->ran_diff;
    if (e->post_diff) {
        if ((st = ensure(e, e->cd_pair9, cc * 9 * sizeof(double)))) return st;
        if ((st = ensure(e, e->cd_vs9, cc * 9 * sizeof(double)))) return st;
        if ((st = ensure(e, e->cd_vo9, cc * 9 * sizeof(double)))) return st;
    }
    CU(e, cudaEventRecord(e->ev_post0, sq));
    const unsigned long long* cnt = (const unsigned long long*)e->rec_count.p;
    const int gx = B >= 8 ? 16 : 148 * 2;
    bh_keys_kernel<<<dim3(gx, B), 256, 0, sq>>>(cnt, e->rec_cap, (const double*)e->rec_p.p, (unsigned long long*)e->sort_keys[0].p,
                                                (unsigned*)e->sort_vals[0].p);
    CU(e, cudaGetLastError());
    RsSegments sg = {nullptr, cnt, e->rec_cap, e->rec_cap};
    if ((st = radix_sort(e, sg, B, max_found, 64, sq))) return st;
    const unsigned long long* keys = (const unsigned long long*)e->sort_keys[0].p;
    bh_tilemin_kernel<<<dim3(ntiles, B), RS_THREADS, 0, sq>>>(keys, cnt, e->rec_cap, ntiles, (double*)e->bh_tmin.p);
    bh_suffix_kernel<<<(B + 63) / 64, 64, 0, sq>>>(B, ntiles, (double*)e->bh_tmin.p);
    bh_q_kernel<<<dim3(ntiles, B), RS_THREADS, 0, sq>>>(keys, (const unsigned*)e->sort_vals[0].p, cnt, e->rec_cap, ntiles,
                                                         (const double*)e->bh_tmin.p, (double*)e->rec_q.p);
    CU(e, cudaGetLastError());
    CU(e, cudaMemsetAsync(e->slotmap.p, 0xFF, (size_t)B * tile * sizeof(int), sq));
    CU(e, cudaMemsetAsync(e->cd_count.p, 0, sizeof(unsigned long long), sq));
    post_slotmap_kernel<<<dim3(gx, B), 256, 0, sq>>>(cnt, e->rec_cap, (const int*)e->rec_row.p, (const int*)e->rec_col.p, e->n, e->wc,
                                                     (int*)e->slotmap.p);
    PostOut po = {(int*)e->cd_block.p, (int*)e->cd_row.p, (int*)e->cd_col.p, (int*)e->cd_flags.p, (double*)e->cd_q.p,
                  (double*)e->cd_sigma.p, (double*)e->cd_cval.p, (double*)e->cd_o9.p, (double*)e->cd_so9.p,
                  e->post_diff ? (double*)e->cd_pair9.p : nullptr, e->post_diff ? (double*)e->cd_vs9.p : nullptr,
                  e->post_diff ? (double*)e->cd_vo9.p : nullptr, (unsigned long long*)e->cd_count.p, e->cand_cap};
    post_select_kernel<<<dim3(gx, B), 256, 0, sq>>>(cnt, e->rec_cap, (const double*)e->rec_q.p, pt, po, (int*)e->cd_slot.p);
    post_candidates_kernel<<<148 * 4, 256, 0, sq>>>(
        e->rec_cap, (const int*)e->rec_row.p, (const int*)e->rec_col.p, (const double*)e->rec_q.p, (const double*)e->rec_sigma.p,
        (const double*)e->rec_v.p, (const double*)e->rec_pair.p, raw_slot(e, e->slot_run), (const int*)e->slotmap.p,
        (const int*)e->cd_slot.p, e->n, e->wc, e->dhi, e->dpx, st_thr, po);
    CU(e, cudaGetLastError());
    CU(e, cudaEventRecord(e->ev_post1, sq));
    CU(e, cudaEventRecord(e->ev_run[e->slot_run], sq));          // the candidate kernel reads the tile slot too
    CU(e, cudaEventRecord(e->ev_post_done, sq));
    e->launches += 7;
    e->post_done = true;
    return MB200_OK;
}

// Enrichment filter of the selected candidates on the device (mb_post.cuh): flags bit 1 = c[x, y] > 2 * mean of the
// non-zero entries of its diagonal (mustache.py:816-828), np.mean reproduced bit for bit; bit 2 = decided.
int mb200_enrich_candidates(mb200_engine* e) {
    if (!e) return MB200_ERR_ARG;
    if (!e->post_done) return fail(e, MB200_ERR_ARG, "mb200_select_candidates has not been called for this batch");
    int st = use_device(e);
    if (st) return st;
    cudaStream_t sq = e->post_stream;
    unsigned long long tot = 0;
    CU(e, cudaMemcpyAsync(&tot, e->cd_count.p, sizeof(tot), cudaMemcpyDeviceToHost, sq));
    CU(e, cudaStreamSynchronize(sq));
    if (tot == 0) return MB200_OK;
    if ((long long)tot > e->cand_cap)
        return fail(e, MB200_ERR_CAPACITY, "%llu candidates, capacity %lld: raise candidate_fraction and call mb200_select_candidates again", tot, e->cand_cap);
    const int nlines = e->nblocks * e->wc;
    // scratch for the compacted diagonals; MB200_ENRICH_POOL_KB shrinks it (tests: forces several rounds)
    const char* pool_kb = std::getenv("MB200_ENRICH_POOL_KB");
    const size_t pool = pool_kb ? (size_t)std::max(1, atoi(pool_kb)) << 10 : (size_t)256 << 20;
    const unsigned max_slots = (unsigned)std::max<size_t>(1, std::min<size_t>(std::min<size_t>(tot, (size_t)nlines),
                                                                             pool / ((size_t)e->n * sizeof(double))));
    if ((st = ensure(e, e->en_need, (size_t)nlines * sizeof(int)))) return st;
    if ((st = ensure(e, e->en_lines, (size_t)max_slots * sizeof(int)))) return st;
    if ((st = ensure(e, e->en_nlist, sizeof(unsigned)))) return st;
    if ((st = ensure(e, e->en_mean, (size_t)max_slots * sizeof(double)))) return st;
    if ((st = ensure(e, e->en_scratch, (size_t)max_slots * e->n * sizeof(double)))) return st;
    const int gc = (int)std::min<unsigned long long>((tot + 255) / 256, 148 * 8);
    const int gl = std::min((nlines + 255) / 256, 148 * 8);
    CU(e, cudaMemsetAsync(e->en_need.p, 0, (size_t)nlines * sizeof(int), sq));
    enrich_mark_kernel<<<gc, 256, 0, sq>>>((const unsigned long long*)e->cd_count.p, e->cand_cap, (const int*)e->cd_block.p,
                                           (const int*)e->cd_row.p, (const int*)e->cd_col.p, (const int*)e->cd_flags.p, e->wc, e->dpx,
                                           (int*)e->en_need.p);
    for (int round = 0; round < 1 + nlines; ++round) {
        CU(e, cudaMemsetAsync(e->en_nlist.p, 0, sizeof(unsigned), sq));
        enrich_list_kernel<<<gl, 256, 0, sq>>>(nlines, (int*)e->en_need.p, (int*)e->en_lines.p, (unsigned*)e->en_nlist.p, max_slots);
        enrich_mean_kernel<<<(int)std::min<unsigned>((max_slots + 7) / 8, 148 * 8), 256, 0, sq>>>(
            raw_slot(e, e->slot_run), e->n, e->wc, (const int*)e->en_lines.p, (const unsigned*)e->en_nlist.p, max_slots,
            (double*)e->en_scratch.p, (double*)e->en_mean.p);
        enrich_apply_kernel<<<gc, 256, 0, sq>>>((const unsigned long long*)e->cd_count.p, e->cand_cap, (const int*)e->cd_block.p,
                                                (const int*)e->cd_row.p, (const int*)e->cd_col.p, (const double*)e->cd_cval.p, e->wc,
                                                e->dpx, (const int*)e->en_need.p, (const double*)e->en_mean.p, (int*)e->cd_flags.p);
        CU(e, cudaGetLastError());
        e->launches += 3;
        unsigned nl = 0;
        CU(e, cudaMemcpyAsync(&nl, e->en_nlist.p, sizeof(nl), cudaMemcpyDeviceToHost, sq));
        CU(e, cudaStreamSynchronize(sq));
        if (nl <= max_slots) break;
        enrich_next_round_kernel<<<gl, 256, 0, sq>>>(nlines, (int*)e->en_need.p);
    }
    CU(e, cudaEventRecord(e->ev_run[e->slot_run], sq));              // the mean kernel reads the tile slot
    CU(e, cudaEventRecord(e->ev_post_done, sq));
    return MB200_OK;
}

int mb200_fetch_candidates(mb200_engine* e, int64_t capacity, int32_t* block, int32_t* row, int32_t* col, int32_t* flags, double* q,
                           double* sigma, double* cval, double* o9, double* so9, double* pair9, double* vself9, double* vother9,
                           int64_t* n_out) {
    if (!e) return MB200_ERR_ARG;
    if (!e->post_done) return fail(e, MB200_ERR_ARG, "mb200_select_candidates has not been called for this batch");
    int st = use_device(e);
    if (st) return st;
    if ((st = refresh_counts(e))) return st;                      // also reports non-finite tiles / record overflow
    for (int b = 0; b < e->nblocks; ++b) {
        if (e->h_nonfinite[b]) return fail(e, MB200_ERR_NONFINITE, "block %d holds non-finite values", b);
        if ((long long)e->h_rec[b] > e->rec_cap)
            return fail(e, MB200_ERR_CAPACITY, "block %d produced %llu records, capacity %lld", b, e->h_rec[b], e->rec_cap);
    }
    unsigned long long tot = 0;
    CU(e, cudaMemcpyAsync(&tot, e->cd_count.p, sizeof(tot), cudaMemcpyDeviceToHost, e->post_stream));
    CU(e, cudaStreamSynchronize(e->post_stream));
    if (n_out) *n_out = (int64_t)tot;
    if ((long long)tot > e->cand_cap)
        return fail(e, MB200_ERR_CAPACITY, "%llu candidates, capacity %lld: raise candidate_fraction and call mb200_select_candidates again", tot, e->cand_cap);
    if (tot == 0 || capacity <= 0) return MB200_OK;
    if ((int64_t)tot > capacity) return fail(e, MB200_ERR_CAPACITY, "%llu candidates, caller capacity %lld", tot, (long long)capacity);
    const size_t m = (size_t)tot;
    if (block) CU(e, cudaMemcpyAsync(block, e->cd_block.p, m * sizeof(int), cudaMemcpyDeviceToHost, e->post_stream));
    if (row) CU(e, cudaMemcpyAsync(row, e->cd_row.p, m * sizeof(int), cudaMemcpyDeviceToHost, e->post_stream));
    if (col) CU(e, cudaMemcpyAsync(col, e->cd_col.p, m * sizeof(int), cudaMemcpyDeviceToHost, e->post_stream));
    if (flags) CU(e, cudaMemcpyAsync(flags, e->cd_flags.p, m * sizeof(int), cudaMemcpyDeviceToHost, e->post_stream));
    if (q) CU(e, cudaMemcpyAsync(q, e->cd_q.p, m * sizeof(double), cudaMemcpyDeviceToHost, e->post_stream));
    if (sigma) CU(e, cudaMemcpyAsync(sigma, e->cd_sigma.p, m * sizeof(double), cudaMemcpyDeviceToHost, e->post_stream));
    if (cval) CU(e, cudaMemcpyAsync(cval, e->cd_cval.p, m * sizeof(double), cudaMemcpyDeviceToHost, e->post_stream));
    if (o9) CU(e, cudaMemcpyAsync(o9, e->cd_o9.p, m * 9 * sizeof(double), cudaMemcpyDeviceToHost, e->post_stream));
    if (so9) CU(e, cudaMemcpyAsync(so9, e->cd_so9.p, m * 9 * sizeof(double), cudaMemcpyDeviceToHost, e->post_stream));
    if (pair9 || vself9 || vother9) {
        if (!e->post_diff) return fail(e, MB200_ERR_ARG, "the batch was not run with mb200_run_differential");
        if (pair9) CU(e, cudaMemcpyAsync(pair9, e->cd_pair9.p, m * 9 * sizeof(double), cudaMemcpyDeviceToHost, e->post_stream));
        if (vself9) CU(e, cudaMemcpyAsync(vself9, e->cd_vs9.p, m * 9 * sizeof(double), cudaMemcpyDeviceToHost, e->post_stream));
        if (vother9) CU(e, cudaMemcpyAsync(vother9, e->cd_vo9.p, m * 9 * sizeof(double), cudaMemcpyDeviceToHost, e->post_stream));
    }
    CU(e, cudaStreamSynchronize(e->post_stream));
    return MB200_OK;
}

int mb200_fetch_q(mb200_engine* e, int block, int64_t capacity, double* q, int64_t* n_out) {
    int64_t nz = 0, nf = 0;
    int st = mb200_block_counts(e, block, &nz, &nf);
    if (n_out) *n_out = nf;
    if (st) return st;
    if (!e->post_done) return fail(e, MB200_ERR_ARG, "mb200_select_candidates has not been called for this batch");
    const int64_t m = std::min<int64_t>(nf, capacity);
    if (m <= 0) return MB200_OK;
    if (!q) return fail(e, MB200_ERR_ARG, "null output array");
    CU(e, cudaMemcpyAsync(q, (double*)e->rec_q.p + (size_t)block * e->rec_cap, m * sizeof(double), cudaMemcpyDeviceToHost, e->post_stream));
    CU(e, cudaStreamSynchronize(e->post_stream));
    return MB200_OK;
}

int mb200_last_post_ms(mb200_engine* e, float* ms) {
    if (!e || !ms) return MB200_ERR_ARG;
    if (!e->post_done) return fail(e, MB200_ERR_ARG, "mb200_select_candidates has not been called for this batch");
    int st = use_device(e);
    if (st) return st;
    CU(e, cudaEventSynchronize(e->ev_post1));
    CU(e, cudaEventElapsedTime(ms, e->ev_post0, e->ev_post1));
    return MB200_OK;
}

int mb200_set_score_sigmas(mb200_engine* e, const double* sigma, int n_scored) {
    if (!e || !sigma) return MB200_ERR_ARG;
    if (!e->have_prog || n_scored != e->prog.n_scored)
        return fail(e, MB200_ERR_ARG, "mb200_set_score_sigmas: %d values for %d scoring steps", n_scored, e->have_prog ? e->prog.n_scored : -1);
    memset(e->score_sigma, 0, sizeof(e->score_sigma));
    memcpy(e->score_sigma, sigma, (size_t)n_scored * sizeof(double));
    return MB200_OK;
}

int mb200_fetch_sigma(mb200_engine* e, int block, int64_t capacity, double* sigma, int64_t* n_out) {
    int64_t nz = 0, nf = 0;
    int st = mb200_block_counts(e, block, &nz, &nf);
    if (n_out) *n_out = nf;
    if (st) return st;
    const int64_t m = std::min<int64_t>(nf, capacity);
    if (m <= 0) return MB200_OK;
    if (!sigma) return fail(e, MB200_ERR_ARG, "null output array");
    CU(e, cudaMemcpyAsync(sigma, (double*)e->rec_sigma.p + (size_t)block * e->rec_cap, m * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaStreamSynchronize(e->stream));
    return MB200_OK;
}

int mb200_records_device(mb200_engine* e, int block, void** rows, void** cols, void** v, void** scored_index, void** p,
                         int64_t* capacity) {
    int st = check_block(e, block);
    if (st) return st;
    if (!e->ran) return fail(e, MB200_ERR_ARG, "mb200_run has not been called for this batch");
    const size_t o = (size_t)block * e->rec_cap;
    if (rows) *rows = (int*)e->rec_row.p + o;
    if (cols) *cols = (int*)e->rec_col.p + o;
    if (v) *v = (double*)e->rec_v.p + o;
    if (scored_index) *scored_index = (int*)e->rec_sidx.p + o;
    if (p) *p = (double*)e->rec_p.p + o;
    if (capacity) *capacity = e->rec_cap;
    return MB200_OK;
}

int mb200_fetch_fits(mb200_engine* e, int block, double* loc, double* scale, int32_t* score_id, int capacity, int* n_scored) {
    int st = check_block(e, block);
    if (st) return st;
    if (!e->ran) return fail(e, MB200_ERR_ARG, "mb200_run has not been called for this batch");
    if (n_scored) *n_scored = e->prog.n_scored;
    const int m = std::min(capacity, e->prog.n_scored);
    if (m <= 0) return MB200_OK;
    if ((st = use_device(e))) return st;
    const size_t o = (size_t)block * e->prog.n_scored;
    if (loc) CU(e, cudaMemcpyAsync(loc, (double*)e->fit_loc.p + o, m * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    if (scale) CU(e, cudaMemcpyAsync(scale, (double*)e->fit_scale.p + o, m * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaStreamSynchronize(e->stream));
    if (score_id)
        for (int t = 0; t < m; ++t) score_id[t] = e->prog.score_id[t];
    return MB200_OK;
}

int mb200_last_timing(mb200_engine* e, float* prep_ms, float* kv_ms, float* kh_ms, float* ks_ms, float* fin_ms,
                      float* total_ms) {
    if (!e) return MB200_ERR_ARG;
    if (!e->ran) return fail(e, MB200_ERR_ARG, "mb200_run has not been called for this batch");
    int st = use_device(e);
    if (st) return st;
    CU(e, cudaEventSynchronize(e->ev_end));
    const int npass = e->npass_run;
    float kv = 0, kh = 0, ks = 0, t = 0;
    for (int p = 0; p < npass; ++p) {
        CU(e, cudaEventElapsedTime(&t, e->ev_pass[4 * p], e->ev_pass[4 * p + 1]));
        kv += t;
        CU(e, cudaEventElapsedTime(&t, e->ev_pass[4 * p + 1], e->ev_pass[4 * p + 2]));
        kh += t;
        CU(e, cudaEventElapsedTime(&t, e->ev_pass[4 * p + 2], e->ev_pass[4 * p + 3]));
        ks += t;
    }
    // a differential run ends with the difference stack (its kernels are reported with the statistics phase)
    cudaEvent_t last = e->ran_diff ? e->ev_diff : e->ev_end;
    CU(e, cudaEventSynchronize(last));
    CU(e, cudaEventElapsedTime(&e->t_prep, e->ev_begin, e->ev_prep));
    CU(e, cudaEventElapsedTime(&e->t_fin, e->ev_pass[4 * (npass - 1) + 3], last));
    e->t_ks = ks;
    if (ks_ms) *ks_ms = ks;
    CU(e, cudaEventElapsedTime(&e->t_total, e->ev_begin, last));
    e->t_kv = kv;
    e->t_kh = kh;
    if (prep_ms) *prep_ms = e->t_prep;
    if (kv_ms) *kv_ms = kv;
    if (kh_ms) *kh_ms = kh;
    if (fin_ms) *fin_ms = e->t_fin;
    if (total_ms) *total_ms = e->t_total;
    return MB200_OK;
}

int mb200_last_launches(mb200_engine* e, int* launches) {
    if (!e || !launches) return MB200_ERR_ARG;
    *launches = e->launches;
    return MB200_OK;
}

int mb200_debug_level(mb200_engine* e, int block, int step, double* gauss_out, double* dog_out) {
    int st = check_block(e, block);
    if (st) return st;
    if (step < 0 || step >= e->prog.n_steps) return fail(e, MB200_ERR_ARG, "step %d out of range", step);
    if ((st = use_device(e))) return st;
    const size_t bytes = (size_t)e->n * e->n * sizeof(double);
    if ((st = ensure(e, e->dbgG, bytes))) return st;
    if ((st = ensure(e, e->dbgL, bytes))) return st;
    CU(e, cudaStreamWaitEvent(e->stream, e->ev_post_done, 0));
    if ((st = adopt_uploads(e))) return st;
    CU(e, cudaMemsetAsync(e->dbgG.p, 0, bytes, e->stream));
    CU(e, cudaMemsetAsync(e->dbgL.p, 0, bytes, e->stream));
    MbGeom g = make_geom(e, block, 1);
    g.dbg_step = step;
    g.dbgG = (double*)e->dbgG.p;
    g.dbgL = (double*)e->dbgL.p;
    // scratch counters of this block are clobbered: the batch has to be re-run before fetching records again
    CU(e, cudaMemsetAsync(g.rec_count, 0, sizeof(unsigned long long), e->stream));
    if ((st = launch_pass(e, block, 1, &g, nullptr))) return st;
    e->ran = false;
    e->counts_valid = false;
    if (gauss_out) CU(e, cudaMemcpyAsync(gauss_out, e->dbgG.p, bytes, cudaMemcpyDeviceToHost, e->stream));
    if (dog_out) CU(e, cudaMemcpyAsync(dog_out, e->dbgL.p, bytes, cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaStreamSynchronize(e->stream));
    return MB200_OK;
}

int mb200_run_differential(mb200_engine* e) {
    if (!e) return MB200_ERR_ARG;
    if (!e->configured) return fail(e, MB200_ERR_ARG, "mb200_configure has not been called");
    if (!e->have_dprog) return fail(e, MB200_ERR_ARG, "mb200_set_diff_program has not been called");
    if (e->nblocks % 2) return fail(e, MB200_ERR_ARG, "differential batches hold map 1 / map 2 tiles in blocks 2k / 2k+1");
    if (e->dprog.rmax > e->prog.rmax) return fail(e, MB200_ERR_ARG, "difference chain radius exceeds the main chain's");
    int st = mb200_run(e);                       // both maps, scored independently (diff_mustache.py:289-425)
    if (st) return st;
    const int npairs = e->nblocks / 2;
    int ndiff = 0, oct_max = 0;
    for (int s = 0; s < e->dprog.n_steps; ++s) ndiff += (e->dprog.st[s].flags & MB_FLAG_DIFFREF) ? 1 : 0;
    for (int t = 0; t < e->prog.n_scored; ++t) oct_max = std::max(oct_max, e->prog.score_id[t] / 12);
    if (oct_max >= ndiff) return fail(e, MB200_ERR_ARG, "difference chain has %d octaves, main chain scores octave %d", ndiff, oct_max);
    const size_t tile = (size_t)e->n * e->wc;
    if (ndiff != e->ndiff) return fail(e, MB200_ERR_ARG, "difference chain changed after mb200_configure");
    if ((st = ensure(e, e->rawD, (size_t)npairs * tile * sizeof(double)))) return st;
    if ((st = ensure(e, e->dout, (size_t)ndiff * npairs * tile * sizeof(double)))) return st;
    if ((st = ensure(e, e->dmu, (size_t)ndiff * npairs * sizeof(double)))) return st;
    if ((st = ensure(e, e->dsd, (size_t)ndiff * npairs * sizeof(double)))) return st;
    if ((st = ensure(e, e->rec_pair, (size_t)e->nblocks * e->rec_cap * sizeof(double)))) return st;
    diff_tile_kernel<<<dim3(148 * 2, npairs), 256, 0, e->stream>>>(raw_slot(e, e->slot_run), (double*)e->rawD.p, e->n, e->wc, e->dpx);
    CU(e, cudaGetLastError());
    CU(e, cudaMemsetAsync(e->dout.p, 0, (size_t)ndiff * npairs * tile * sizeof(double), e->stream));
    // difference stack: same kernels, constant regions are 0 (c = zeros; c[nz] = c1[nz] - c2[nz]), nothing is scored.
    // dout is [pair][octave][n][wc], so every pass of pass_blocks pairs writes its own slice.
    for (int first = 0; first < npairs; first += e->pass_blocks) {
        const int nb = std::min(e->pass_blocks, npairs - first);
        MbGeom g = make_geom(e, 0, nb);
        g.raw = (const double*)e->rawD.p + (size_t)first * tile;
        g.fill = 0.0;
        g.rec_cap = 0;
        g.L = nullptr;                      // nothing is scored on the difference stack: its DoGs go to dout only
        g.dout = (double*)e->dout.p + (size_t)first * ndiff * tile;
        if ((st = launch_pass(e, 0, nb, &g, nullptr, &e->dprog))) return st;
    }
    {
        const int items = ndiff * npairs;
        if ((st = ensure(e, e->dpart, (size_t)items * (DIFF_NCH * 2 + 1) * sizeof(double)))) return st;
        double* part = (double*)e->dpart.p;
        double* cntv = part + (size_t)items * DIFF_NCH * 2;
        const dim3 gp(DIFF_NCH, ndiff, npairs);
        diff_partial_kernel<0><<<gp, 256, 0, e->stream>>>(raw_slot(e, e->slot_run), (const double*)e->dout.p, e->n, e->wc, nullptr, part);
        diff_finish_kernel<0><<<(items + 63) / 64, 64, 0, e->stream>>>(part, items, (double*)e->dmu.p, (double*)e->dsd.p, cntv);
        diff_partial_kernel<1><<<gp, 256, 0, e->stream>>>(raw_slot(e, e->slot_run), (const double*)e->dout.p, e->n, e->wc,
                                                          (const double*)e->dmu.p, part);
        diff_finish_kernel<1><<<(items + 63) / 64, 64, 0, e->stream>>>(part, items, (double*)e->dmu.p, (double*)e->dsd.p, cntv);
        CU(e, cudaGetLastError());
        e->launches += 3;
    }
    diff_pair_kernel<<<dim3(32, e->nblocks), 256, 0, e->stream>>>(
        (const unsigned long long*)e->rec_count.p, e->rec_cap, (const int*)e->rec_row.p, (const int*)e->rec_col.p,
        (const int*)e->rec_sidx.p, (const int*)e->d_score_id.p, (const double*)e->dout.p, (const double*)e->dmu.p,
        (const double*)e->dsd.p, e->n, e->wc, ndiff, (double*)e->rec_pair.p);
    CU(e, cudaGetLastError());
    e->launches += 3;
    CU(e, cudaEventRecord(e->ev_diff, e->stream));
    CU(e, cudaEventRecord(e->ev_run[e->slot_run], e->stream));   // the difference kernels read the tile slot too
    e->ran_diff = true;
    return MB200_OK;
}

int mb200_fetch_pair(mb200_engine* e, int block, int64_t capacity, double* pair, int64_t* n_out) {
    int64_t nz = 0, nf = 0;
    int st = mb200_block_counts(e, block, &nz, &nf);
    if (n_out) *n_out = nf;
    if (st) return st;
    if (!e->ran_diff) return fail(e, MB200_ERR_ARG, "mb200_run_differential has not been called for this batch");
    const int64_t m = std::min<int64_t>(nf, capacity);
    if (m <= 0) return MB200_OK;
    if (!pair) return fail(e, MB200_ERR_ARG, "null output array");
    CU(e, cudaMemcpyAsync(pair, (double*)e->rec_pair.p + (size_t)block * e->rec_cap, m * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CU(e, cudaStreamSynchronize(e->stream));
    return MB200_OK;
}

int mb200_normalize_sparse(mb200_engine* e, const int32_t* x, const int32_t* y, double* v, int64_t nnz, int resolution,
                           int distance_in_px, double* weights, int weights_cap, int* n_weights) {
    if (!e) return MB200_ERR_ARG;
    if (n_weights) *n_weights = 0;
    if (nnz < 0 || (nnz > 0 && (!x || !y || !v)) || resolution < 1 || distance_in_px < 0)
        return fail(e, MB200_ERR_ARG, "bad arguments to mb200_normalize_sparse");
    if (nnz == 0) return MB200_OK;
    if (nnz >= (1LL << 32)) return fail(e, MB200_ERR_ARG, "more than 2^32 contacts in one chromosome");
    int st = use_device(e);
    if (st) return st;
    long long n = 0;
    for (int64_t k = 0; k < nnz; ++k) {
        if (x[k] < 0 || y[k] < 0) return fail(e, MB200_ERR_ARG, "negative bin index at contact %lld", (long long)k);
        n = std::max<long long>(n, std::max(x[k], y[k]));
    }
    n += 1;                                                                   // mustache.py:623
    if (n >= (1LL << NZ_POS_BITS)) return fail(e, MB200_ERR_ARG, "chromosome of %lld bins (limit %d)", n, 1 << NZ_POS_BITS);
    const bool windowed = (n - distance_in_px) * (long long)resolution > 2000000;      // mustache.py:628
    const long long D = windowed ? (long long)distance_in_px + 2 : std::min<long long>(distance_in_px, n);
    const size_t m8 = (size_t)nnz * sizeof(double), m4 = (size_t)nnz * sizeof(int);
    if ((st = ensure(e, e->nz_x, m4))) return st;
    if ((st = ensure(e, e->nz_y, m4))) return st;
    if ((st = ensure(e, e->nz_v, m8))) return st;
    if ((st = ensure(e, e->nz_xs, m4))) return st;
    if ((st = ensure(e, e->nz_vs, m8))) return st;
    if ((st = ensure(e, e->nz_out, m8))) return st;
    for (int k = 0; k < 2; ++k) {
        if ((st = ensure(e, e->sort_keys[k], (size_t)nnz * sizeof(unsigned long long)))) return st;
        if ((st = ensure(e, e->sort_vals[k], (size_t)nnz * sizeof(unsigned)))) return st;
    }
    if ((st = ensure(e, e->nz_seg, (D + 2) * sizeof(long long)))) return st;
    if ((st = ensure(e, e->nz_mean, (D + 1) * sizeof(double)))) return st;
    if ((st = ensure(e, e->nz_sd, (D + 1) * sizeof(double)))) return st;
    if ((st = ensure(e, e->nz_w, (D + 1) * sizeof(double)))) return st;
    cudaStream_t sq = e->stream;
    e->launches = 0;
    CU(e, cudaMemcpyAsync(e->nz_x.p, x, m4, cudaMemcpyHostToDevice, sq));
    CU(e, cudaMemcpyAsync(e->nz_y.p, y, m4, cudaMemcpyHostToDevice, sq));
    CU(e, cudaMemcpyAsync(e->nz_v.p, v, m8, cudaMemcpyHostToDevice, sq));
    const int grid = (int)std::min<long long>((nnz + 255) / 256, 148LL * 16);
    // 1. contacts grouped by diagonal, by position inside a diagonal (stable: input order for row-sorted input);
    //    contacts the reference's loop never visits (|y - x| >= D) land in bucket D and pass through
    nz_keys_kernel<<<grid, 256, 0, sq>>>((const int*)e->nz_x.p, (const int*)e->nz_y.p, nnz, (int)D,
                                         (unsigned long long*)e->sort_keys[0].p, (unsigned*)e->sort_vals[0].p);
    CU(e, cudaGetLastError());
    std::vector<long long> one_seg = {0, (long long)nnz};
    CU(e, cudaMemcpyAsync(e->nz_seg.p, one_seg.data(), 2 * sizeof(long long), cudaMemcpyHostToDevice, sq));
    CU(e, cudaStreamSynchronize(sq));                                          // one_seg is a local
    RsSegments sg = {(const long long*)e->nz_seg.p, nullptr, 0, 0};
    int dbits = 1;
    while ((1LL << dbits) <= D) ++dbits;
    // the sort reads its single segment's bounds from nz_seg[0..1]; the diagonals' bounds replace them afterwards
    if ((st = radix_sort(e, sg, 1, nnz, NZ_POS_BITS + dbits, sq))) return st;
    const unsigned long long* keys = (const unsigned long long*)e->sort_keys[0].p;
    const unsigned* idx = (const unsigned*)e->sort_vals[0].p;
    nz_segments_kernel<<<(unsigned)((D + 2 + 255) / 256), 256, 0, sq>>>(keys, nnz, (int)D, (long long*)e->nz_seg.p);
    nz_gather_kernel<<<grid, 256, 0, sq>>>(keys, idx, (const double*)e->nz_v.p, nnz, windowed ? 0 : 1, (int*)e->nz_xs.p,
                                           (double*)e->nz_vs.p);
    CU(e, cudaGetLastError());
    // 2. np.mean / np.std per diagonal
    if (D > 0) {
        nz_stats_kernel<<<(unsigned)((D * 8 + 255) / 256), 256, 0, sq>>>((const double*)e->nz_vs.p, (const long long*)e->nz_seg.p,
                                                                         (int)D, (double*)e->nz_mean.p, (double*)e->nz_sd.p);
        CU(e, cudaGetLastError());
    }
    e->launches += 4;
    // 3. z-scores; the pass-through bucket keeps its (cleaned) values
    if (windowed) {
        std::vector<double> mean(D), w(D);
        CU(e, cudaMemcpyAsync(mean.data(), e->nz_mean.p, D * sizeof(double), cudaMemcpyDeviceToHost, sq));
        CU(e, cudaStreamSynchronize(sq));
        // 1 + math.log(1 + mean, 30) with the host's libm, as the reference evaluates it (mustache.py:667-668)
        for (long long d = 0; d < D; ++d) w[d] = 1.0 + std::log(1.0 + mean[d]) / std::log(30.0);
        CU(e, cudaMemcpyAsync(e->nz_w.p, w.data(), D * sizeof(double), cudaMemcpyHostToDevice, sq));
        if (weights)
            for (long long d = 0; d < D && d < weights_cap; ++d) weights[d] = w[d];
        if (n_weights) *n_weights = (int)D;
        CU(e, cudaMemcpyAsync(e->nz_out.p, e->nz_v.p, m8, cudaMemcpyDeviceToDevice, sq));      // pass-through default
        long long h_seg_D = 0;
        CU(e, cudaMemcpyAsync(&h_seg_D, (long long*)e->nz_seg.p + D, sizeof(long long), cudaMemcpyDeviceToHost, sq));
        CU(e, cudaStreamSynchronize(sq));                                      // w is a local; h_seg_D = visited contacts
        if (h_seg_D > 0) {
            const int wgrid = (int)std::min<long long>((h_seg_D + 7) / 8, 148LL * 32);
            nz_window_kernel<<<wgrid, 256, 0, sq>>>(keys, (const int*)e->nz_xs.p, (const double*)e->nz_vs.p, idx,
                                                    (const long long*)e->nz_seg.p, h_seg_D, n, (int)(2000000 / resolution),
                                                    (const double*)e->nz_mean.p, (const double*)e->nz_sd.p,
                                                    (const double*)e->nz_w.p, (double*)e->nz_out.p);
            CU(e, cudaGetLastError());
            e->launches += 1;
        }
    } else {
        nz_global_kernel<<<grid, 256, 0, sq>>>(keys, idx, (const double*)e->nz_vs.p, nnz, (int)D, (const double*)e->nz_mean.p,
                                               (const double*)e->nz_sd.p, (double*)e->nz_out.p);
        CU(e, cudaGetLastError());
        e->launches += 1;
    }
    CU(e, cudaMemcpyAsync(v, e->nz_out.p, m8, cudaMemcpyDeviceToHost, sq));
    CU(e, cudaStreamSynchronize(sq));
    return MB200_OK;
}

int mb200_kv_plan(int n_steps, const int32_t* radius, int32_t* group_of_step, int64_t* fp64_per_output) {
    if (!radius || n_steps < 1 || n_steps > MB_MAX_STEPS) return MB200_ERR_ARG;
    static MbProgram p;                      // host-only helper, no engine: large structs stay off the stack
    static KvPlan kp;
    memset(&p, 0, sizeof(p));
    p.n_steps = n_steps;
    int off = 0;
    for (int s = 0; s < n_steps; ++s) {
        if (radius[s] < 1 || off + radius[s] + 1 > MB_MAX_TAPS) return MB200_ERR_ARG;
        p.st[s].radius = radius[s];
        p.st[s].tap_off = off;
        off += radius[s] + 1;
        p.rmax = std::max(p.rmax, (int)radius[s]);
    }
    int st = plan_kv(nullptr, p, kp);
    if (st) return st;
    long long cost = 0;
    for (int gi = 0; gi < kp.n_groups; ++gi) {
        const KvGroup& gr = kp.grp[gi];
        cost += (long long)gr.rmax * (2 * gr.n + 1) + gr.n;
        for (int slot = 0; slot < gr.n; ++slot)
            if (group_of_step) group_of_step[gr.step[slot]] = gi;
    }
    if (fp64_per_output) *fp64_per_output = cost;
    return MB200_OK;
}

int mb200_kh_ring_plan(int n_steps, const int32_t* radius, int32_t* offset, int32_t* size, int32_t* dep, int32_t* ring_doubles) {
    if (!radius || n_steps < 1 || n_steps > MB_MAX_STEPS) return MB200_ERR_ARG;
    static MbProgram p;                      // host-only helper, no engine: large structs stay off the stack
    memset(&p, 0, sizeof(p));
    p.n_steps = n_steps;
    for (int s = 0; s < n_steps; ++s) {
        if (radius[s] < 1) return MB200_ERR_ARG;
        p.st[s].radius = radius[s];
        p.rmax = std::max(p.rmax, (int)radius[s]);
    }
    if (kh_smem_bytes(p.rmax, 1) > 227 * 1024) return MB200_ERR_ARG;
    plan_kh_ring(p);
    for (int s = 0; s < n_steps; ++s) {
        if (offset) offset[s] = p.stage[s].off;
        if (size) size[s] = (KH_TR * kh_box_width(p.st[s].radius) + 15) & ~15;
        if (dep) dep[s] = p.stage[s].dep;
    }
    if (ring_doubles) *ring_doubles = kh_ring_doubles(p.rmax);
    return MB200_OK;
}

int mb200_contacts_open(const char* path, const char* chromosome, int threads, void** handle, int64_t* n_rows, int* n_cols,
                        int* value_is_int) {
    if (!path || !handle) return MB200_ERR_ARG;
    *handle = nullptr;
    mb200_contacts* c = new mb200_contacts();
    const int rc = mb200_parse_file(path, chromosome, threads, c);
    if (rc != 0) {
        delete c;
        return rc < 0 ? MB200_ERR_ARG : rc;             // MB200_PARSE_UNSUPPORTED (1): the caller keeps its pandas reader
    }
    *handle = c;
    if (n_rows) *n_rows = (int64_t)c->a.size();
    if (n_cols) *n_cols = c->ncols;
    if (value_is_int) *value_is_int = c->value_is_int ? 1 : 0;
    return MB200_OK;
}

int mb200_contacts_read(void* handle, int64_t* a, int64_t* b, double* val) {
    if (!handle || !a || !b || !val) return MB200_ERR_ARG;
    const mb200_contacts* c = (const mb200_contacts*)handle;
    if (!c->a.empty()) {
        memcpy(a, c->a.data(), c->a.size() * sizeof(int64_t));
        memcpy(b, c->b.data(), c->b.size() * sizeof(int64_t));
        memcpy(val, c->val.data(), c->val.size() * sizeof(double));
    }
    return MB200_OK;
}

void mb200_contacts_close(void* handle) { delete (mb200_contacts*)handle; }

int mb200_host_alloc(void** ptr, int64_t bytes) {
    if (!ptr || bytes <= 0) return MB200_ERR_ARG;
    return cudaHostAlloc(ptr, (size_t)bytes, cudaHostAllocDefault) == cudaSuccess ? MB200_OK : MB200_ERR_NOMEM;
}

int mb200_host_free(void* ptr) { return cudaFreeHost(ptr) == cudaSuccess ? MB200_OK : MB200_ERR_CUDA; }

int mb200_scale_space_dense(mb200_engine* e, const double* tile, int n, int64_t ld, int dpx, int intra, int64_t capacity,
                            int32_t* rows, int32_t* cols, double* v, int32_t* score_id, double* p, int64_t* nz_count,
                            int64_t* n_found) {
    int st = mb200_configure(e, n, dpx, intra, 1, -1.0);
    if (st) return st;
    if ((st = mb200_upload_dense_host(e, 0, tile, ld))) return st;
    if ((st = mb200_run(e))) return st;
    int64_t nz = 0, nf = 0;
    st = mb200_block_counts(e, 0, &nz, &nf);
    if (nz_count) *nz_count = nz;
    if (n_found) *n_found = nf;
    if (st) return st;
    if (nf > capacity) return fail(e, MB200_ERR_CAPACITY, "%lld records, caller capacity %lld", (long long)nf, (long long)capacity);
    return mb200_fetch_records(e, 0, capacity, rows, cols, v, score_id, p, nullptr);
}

}  // extern "C"
