"""ctypes binding of the C-ABI scale-space engine (include/mustache_b200.h) and a small Python wrapper.

There is NO CPU fallback: if the CUDA library is missing or no device is visible, construction raises.
"""
import ctypes as C
import os

import numpy as np

from . import ladder

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libmustache_b200.so")

ERR_NAMES = {-1: "MB200_ERR_CUDA", -2: "MB200_ERR_ARG", -3: "MB200_ERR_CAPACITY", -4: "MB200_ERR_NONFINITE",
             -5: "MB200_ERR_NOMEM"}
STEP_RESTART, STEP_SCORE, STEP_DIFFREF = 1, 2, 4

# every symbol include/mustache_b200.h declares: name -> (restype, argtypes)
_i32p, _f64p, _i64p = C.POINTER(C.c_int32), C.POINTER(C.c_double), C.POINTER(C.c_int64)
_f32p = C.POINTER(C.c_float)
_H = C.c_void_p
SIGNATURES = {
    "mb200_abi_version": (C.c_int, []),
    "mb200_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "mb200_create": (C.c_int, [C.c_int, C.POINTER(_H)]),
    "mb200_destroy": (None, [_H]),
    "mb200_last_error": (C.c_char_p, [_H]),
    "mb200_set_program": (C.c_int, [_H, C.c_int, _i32p, _i32p, _i32p, _i32p, _f64p, C.c_int]),
    "mb200_configure": (C.c_int, [_H, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double]),
    "mb200_upload_coo_host": (C.c_int, [_H, C.c_int, _i32p, _i32p, _f64p, C.c_int64]),
    "mb200_upload_coo_batch": (C.c_int, [_H, C.c_int, C.c_int, _i64p, _i32p, _i32p, _f64p]),
    "mb200_upload_coo_dev": (C.c_int, [_H, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]),
    "mb200_upload_dense_host": (C.c_int, [_H, C.c_int, C.c_void_p, C.c_int64]),
    "mb200_upload_dense_dev": (C.c_int, [_H, C.c_int, C.c_void_p, C.c_int64]),
    "mb200_upload_band_host": (C.c_int, [_H, C.c_int, C.c_void_p, C.c_int64]),
    "mb200_run": (C.c_int, [_H]),
    "mb200_sync": (C.c_int, [_H]),
    "mb200_run_after": (C.c_int, [_H, _H]),
    "mb200_block_counts": (C.c_int, [_H, C.c_int, _i64p, _i64p]),
    "mb200_fetch_records": (C.c_int, [_H, C.c_int, C.c_int64, _i32p, _i32p, _f64p, _i32p, _f64p, _i64p]),
    "mb200_batch_counts": (C.c_int, [_H, _i64p, _i64p]),
    "mb200_pack_records": (C.c_int, [_H, _i64p, _i64p]),
    "mb200_fetch_packed": (C.c_int, [_H, C.c_int64, _i32p, _i32p, _f64p, _i32p, _f64p, _f64p, _f64p]),
    "mb200_packed_device": (C.c_int, [_H] + [C.POINTER(C.c_void_p)] * 8),
    "mb200_select_candidates": (C.c_int, [_H, C.c_double, C.c_double, C.c_double]),
    "mb200_enrich_candidates": (C.c_int, [_H]),
    "mb200_fetch_candidates": (C.c_int, [_H, C.c_int64, _i32p, _i32p, _i32p, _i32p, _f64p, _f64p, _f64p, _f64p, _f64p, _f64p,
                                         _f64p, _f64p, _i64p]),
    "mb200_fetch_q": (C.c_int, [_H, C.c_int, C.c_int64, _f64p, _i64p]),
    "mb200_last_post_ms": (C.c_int, [_H, _f32p]),
    "mb200_set_arithmetic": (C.c_int, [_H, C.c_int]),
    "mb200_set_fusion": (C.c_int, [_H, C.c_int]),
    "mb200_set_overlap": (C.c_int, [_H, C.c_int]),
    "mb200_set_pass_limit": (C.c_int, [_H, C.c_int]),
    "mb200_set_score_sigmas": (C.c_int, [_H, _f64p, C.c_int]),
    "mb200_fetch_sigma": (C.c_int, [_H, C.c_int, C.c_int64, _f64p, _i64p]),
    "mb200_records_device": (C.c_int, [_H, C.c_int] + [C.POINTER(C.c_void_p)] * 5 + [_i64p]),
    "mb200_fetch_fits": (C.c_int, [_H, C.c_int, _f64p, _f64p, _i32p, C.c_int, C.POINTER(C.c_int)]),
    "mb200_last_timing": (C.c_int, [_H, _f32p, _f32p, _f32p, _f32p, _f32p, _f32p]),
    "mb200_last_launches": (C.c_int, [_H, C.POINTER(C.c_int)]),
    "mb200_debug_level": (C.c_int, [_H, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "mb200_set_diff_program": (C.c_int, [_H, C.c_int, _i32p, _i32p, _i32p, _f64p, C.c_int]),
    "mb200_run_differential": (C.c_int, [_H]),
    "mb200_fetch_pair": (C.c_int, [_H, C.c_int, C.c_int64, _f64p, _i64p]),
    "mb200_normalize_sparse": (C.c_int, [_H, _i32p, _i32p, _f64p, C.c_int64, C.c_int, C.c_int, _f64p, C.c_int,
                                         C.POINTER(C.c_int)]),
    "mb200_kv_plan": (C.c_int, [C.c_int, _i32p, _i32p, _i64p]),
    "mb200_kh_ring_plan": (C.c_int, [C.c_int, _i32p, _i32p, _i32p, _i32p, _i32p]),
    "mb200_contacts_open": (C.c_int, [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_void_p), _i64p, C.POINTER(C.c_int),
                                      C.POINTER(C.c_int)]),
    "mb200_contacts_read": (C.c_int, [C.c_void_p, _i64p, _i64p, _f64p]),
    "mb200_contacts_close": (None, [C.c_void_p]),
    "mb200_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64]),
    "mb200_host_free": (C.c_int, [C.c_void_p]),
    "mb200_scale_space_dense": (C.c_int, [_H, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int64, _i32p, _i32p,
                                          _f64p, _i32p, _f64p, _i64p, _i64p]),
}

_lib = None


class EngineError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("%s (%d): %s" % (ERR_NAMES.get(code, "MB200_ERR"), code, msg))
        self.code = code


def load_library(path=None):
    """dlopen the CUDA library and bind every declared symbol.  Raises if the library has not been built."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("MUSTACHE_B200_LIB") or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError("CUDA library %s not built (run `python -m mustache_b200.build`); there is no CPU fallback" % p)
    lib = C.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def _ptr(a, typ):
    return a.ctypes.data_as(typ)


def program_arrays(prog):
    """ScaleProgram -> the flat arrays mb200_set_program takes."""
    radius, flags, sid, off, taps = [], [], [], [], []
    for st in prog.steps:
        r = st.radius
        half = np.ascontiguousarray(st.taps[r:])            # w[0..R], centre first (taps are exactly symmetric)
        assert np.array_equal(half, st.taps[:r + 1][::-1])
        off.append(len(taps))
        taps.extend(half.tolist())
        radius.append(r)
        flags.append((STEP_RESTART if st.restart else 0) | (STEP_SCORE if st.score_id else 0)
                     | (STEP_DIFFREF if st.diff_ref else 0))
        sid.append(st.score_id)
    return (np.array(radius, np.int32), np.array(flags, np.int32), np.array(sid, np.int32), np.array(off, np.int32),
            np.array(taps, np.float64))


class PinnedBuffer:
    """Page-locked host array (mb200_host_alloc) for full-speed uploads."""

    def __init__(self, shape, dtype=np.float64):
        self.lib = load_library()
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        self.ptr = C.c_void_p()
        st = self.lib.mb200_host_alloc(C.byref(self.ptr), nbytes)
        if st:
            raise EngineError(st, "pinned allocation of %d bytes failed" % nbytes)
        buf = (C.c_char * nbytes).from_address(self.ptr.value)
        self.array = np.frombuffer(buf, dtype=dtype).reshape(shape)

    def free(self):
        if self.ptr:
            self.array = None
            self.lib.mb200_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class ScaleSpaceEngine:
    """One engine per GPU.  Typical use:
        eng = ScaleSpaceEngine(0); eng.set_octaves([1.6, 3.2]); eng.configure(n, dpx, nblocks)
        eng.upload_coo(b, rows, cols, vals) ...; eng.run(); rec = eng.records(b)
    """

    def __init__(self, device=0, lib_path=None):
        self.lib = load_library(lib_path)
        cnt = C.c_int(0)
        self.lib.mb200_device_count(C.byref(cnt))
        if cnt.value < 1:
            raise RuntimeError("no CUDA device visible: the scale-space engine has no CPU fallback")
        self.h = _H()
        st = self.lib.mb200_create(int(device), C.byref(self.h))
        if st:
            raise EngineError(st, "mb200_create(device=%d) failed" % device)
        self.device = int(device)
        self._pin = {}
        self._keep = []
        self.program = None
        self.n = self.dpx = self.nblocks = None

    def _pinned(self, name, n, dtype):
        """Engine-owned page-locked staging array (grow-only) for full-speed device -> host record copies."""
        buf = self._pin.get(name)
        if buf is None or buf.array.size < n:
            if buf is not None:
                buf.free()
            buf = PinnedBuffer((max(4096, int(1.25 * n)),), dtype)
            self._pin[name] = buf
        return buf.array[:n]

    def close(self):
        for buf in getattr(self, "_pin", {}).values():
            buf.free()
        self._pin = {}
        if getattr(self, "h", None):
            self.lib.mb200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, st):
        if st:
            raise EngineError(st, self.lib.mb200_last_error(self.h).decode())

    # ---- scale table ----
    def set_program(self, prog):
        radius, flags, sid, off, taps = program_arrays(prog)
        self._chk(self.lib.mb200_set_program(self.h, len(radius), _ptr(radius, _i32p), _ptr(flags, _i32p), _ptr(sid, _i32p),
                                             _ptr(off, _i32p), _ptr(taps, _f64p), len(taps)))
        self.program = prog
        sig = np.array([st.score_sigma for st in prog.steps if st.score_id], np.float64)
        if len(sig):
            self._chk(self.lib.mb200_set_score_sigmas(self.h, _ptr(sig, _f64p), len(sig)))

    def _sigma_lut(self):
        lut = np.zeros(256)
        for k, sg in self.program.sigma_of_id.items():
            lut[k] = sg
        return lut

    def set_octaves(self, octave_values, dedupe=True, differential=False):
        self.set_program(ladder.build_program(list(octave_values), dedupe=dedupe))
        if differential:
            radius, flags, sid, off, taps = program_arrays(ladder.build_diff_program(list(octave_values)))
            self._chk(self.lib.mb200_set_diff_program(self.h, len(radius), _ptr(radius, _i32p), _ptr(flags, _i32p),
                                                      _ptr(off, _i32p), _ptr(taps, _f64p), len(taps)))

    # ---- batch ----
    def configure(self, n, dpx, nblocks=1, intra=True, record_fraction=-1.0):
        self._chk(self.lib.mb200_configure(self.h, int(n), int(dpx), 1 if intra else 0, int(nblocks), float(record_fraction)))
        self.n, self.dpx, self.nblocks = int(n), int(dpx), int(nblocks)

    def upload_coo(self, block, rows, cols, vals):
        rows = np.ascontiguousarray(rows, np.int32)
        cols = np.ascontiguousarray(cols, np.int32)
        vals = np.ascontiguousarray(vals, np.float64)
        self._chk(self.lib.mb200_upload_coo_host(self.h, int(block), _ptr(rows, _i32p), _ptr(cols, _i32p), _ptr(vals, _f64p),
                                                 len(vals)))

    def upload_coo_batch(self, first_block, offsets, rows, cols, vals):
        """Blocks first_block .. first_block + len(offsets) - 2 in one call: concatenated block-local COO (int32, int32,
        float64) with entries [offsets[b], offsets[b+1]) belonging to block first_block + b."""
        offsets = np.ascontiguousarray(offsets, np.int64)
        assert rows.dtype == np.int32 and cols.dtype == np.int32 and vals.dtype == np.float64
        self._chk(self.lib.mb200_upload_coo_batch(self.h, int(first_block), len(offsets) - 1, _ptr(offsets, _i64p),
                                                  _ptr(rows, _i32p), _ptr(cols, _i32p), _ptr(vals, _f64p)))

    def upload_coo_dev(self, block, rows, cols, vals):
        """Block-local COO already on the engine's device (torch tensors: int32, int32, float64), complete on their stream."""
        assert rows.is_cuda and rows.dtype.itemsize == 4 and vals.dtype.itemsize == 8 and rows.is_contiguous()
        self._keep.append((rows, cols, vals))                 # the scatter runs asynchronously: keep the tensors alive
        self._chk(self.lib.mb200_upload_coo_dev(self.h, int(block), C.c_void_p(rows.data_ptr()), C.c_void_p(cols.data_ptr()),
                                                C.c_void_p(vals.data_ptr()), int(vals.numel())))

    def set_arithmetic(self, fused_multiply_add):
        """False: scipy's multiply-then-add (bit-exact Gaussians, default).  True: opt-in FMA fast mode."""
        self._chk(self.lib.mb200_set_arithmetic(self.h, 1 if fused_multiply_add else 0))

    def set_fusion(self, mode):
        """0 / False (default): three kernels.  1 / True: fused axis-1 + scoring kernel (khs_kernel) whenever the chain
        fits.  2: fused axis-0 + axis-1 kernel (kvh_kernel) for chains up to radius 14."""
        self._chk(self.lib.mb200_set_fusion(self.h, int(mode)))

    def set_overlap(self, enable):
        """True: two half-batches in flight on two streams (scoring of one overlaps the Gaussian passes of the other)."""
        self._chk(self.lib.mb200_set_overlap(self.h, 1 if enable else 0))

    def set_pass_limit(self, max_blocks):
        """At most this many blocks per pass of the kernels (0 = whatever fits); next configure()."""
        self._chk(self.lib.mb200_set_pass_limit(self.h, int(max_blocks)))

    def upload_dense(self, block, tile):
        """tile: C-contiguous float64 numpy array (n x n) or anything with .data_ptr() on the engine's device."""
        if hasattr(tile, "data_ptr"):
            assert tile.is_cuda and tile.dtype.is_floating_point and tile.element_size() == 8 and tile.stride(1) == 1
            self._chk(self.lib.mb200_upload_dense_dev(self.h, int(block), C.c_void_p(tile.data_ptr()), int(tile.stride(0))))
            return
        assert tile.dtype == np.float64 and tile.ndim == 2 and tile.strides[1] == 8
        self._keep.append(tile)                               # the copy is asynchronous: the host buffer must outlive it
        self._chk(self.lib.mb200_upload_dense_host(self.h, int(block), C.c_void_p(tile.ctypes.data), tile.strides[0] // 8))

    def upload_band(self, block, band):
        assert band.dtype == np.float64 and band.ndim == 2 and band.strides[1] == 8 and band.shape[0] == self.n
        self._keep.append(band)
        self._chk(self.lib.mb200_upload_band_host(self.h, int(block), C.c_void_p(band.ctypes.data), band.strides[0] // 8))

    def run(self, sync=False):
        self._chk(self.lib.mb200_run(self.h))
        if sync:
            self.sync()

    def run_after(self, other):
        """Everything enqueued on this engine from now on starts after what `other` (same GPU) has enqueued so far: two
        engines used alternately keep the GPU on the runs while the post-processing of the previous batch overlaps."""
        self._chk(self.lib.mb200_run_after(self.h, other.h))

    def run_differential(self):
        """Blocks 2k / 2k+1 = map 1 / map 2 of pair k: both maps scored + pPair of every record."""
        self._chk(self.lib.mb200_run_differential(self.h))

    def sync(self):
        self._chk(self.lib.mb200_sync(self.h))
        self._keep = []

    def counts(self, block):
        nz, nf = C.c_int64(0), C.c_int64(0)
        self._chk(self.lib.mb200_block_counts(self.h, int(block), C.byref(nz), C.byref(nf)))
        return nz.value, nf.value

    def records(self, block, sort=True, pair=False, pinned=False):
        """Records of every updated pixel, sorted row-major (the order of c[nz] in the reference).
        pinned=True: the arrays are views of engine-owned page-locked buffers, valid until the next records() call."""
        nz, nf = self.counts(block)
        pp = None
        if pair:
            pp = np.empty(nf, np.float64)
            n2 = C.c_int64(0)
            self._chk(self.lib.mb200_fetch_pair(self.h, int(block), nf, _ptr(pp, _f64p), C.byref(n2)))
        if pinned:
            rows, cols, sid = (self._pinned(k, nf, np.int32) for k in ("rows", "cols", "sid"))
            v, p, sig = (self._pinned(k, nf, np.float64) for k in ("v", "p", "sigma"))
        else:
            rows, cols = np.empty(nf, np.int32), np.empty(nf, np.int32)
            v, p, sid = np.empty(nf, np.float64), np.empty(nf, np.float64), np.empty(nf, np.int32)
            sig = np.empty(nf, np.float64)
        n_out = C.c_int64(0)
        self._chk(self.lib.mb200_fetch_records(self.h, int(block), nf, _ptr(rows, _i32p), _ptr(cols, _i32p), _ptr(v, _f64p),
                                               _ptr(sid, _i32p), _ptr(p, _f64p), C.byref(n_out)))
        self._chk(self.lib.mb200_fetch_sigma(self.h, int(block), nf, _ptr(sig, _f64p), C.byref(n_out)))   # resolved on the device
        if sort and nf:
            order = np.lexsort((cols, rows))
            rows, cols, v, p, sid, sig = rows[order], cols[order], v[order], p[order], sid[order], sig[order]
            if pp is not None:
                pp = pp[order]
        out = dict(rows=rows, cols=cols, v=v, p=p, score_id=sid, sigma=sig, nz_count=nz, n_found=nf)
        if pp is not None:
            out["pair"] = pp
        return out

    def batch_counts(self):
        nz, nf = np.zeros(self.nblocks, np.int64), np.zeros(self.nblocks, np.int64)
        self._chk(self.lib.mb200_batch_counts(self.h, _ptr(nz, _i64p), _ptr(nf, _i64p)))
        self._keep = []                                       # the run has completed: device inputs may be released
        return nz, nf

    def pack(self):
        """Pack the batch's records block after block on the device; returns the offsets array (nblocks + 1)."""
        off = np.zeros(self.nblocks + 1, np.int64)
        tot = C.c_int64(0)
        self._chk(self.lib.mb200_pack_records(self.h, _ptr(off, _i64p), C.byref(tot)))
        return off

    def records_batch(self, sort=True, pair=False, pinned=True):
        """Records of every block of the batch with ONE device -> host copy per field (page-locked staging).
        Returns a list of per-block dicts like records(); with pinned=True the arrays are views of engine-owned buffers
        that stay valid until the next records_batch() call (sort=True copies them)."""
        off = self.pack()
        tot = int(off[-1])
        nz, _ = self.batch_counts()
        if pinned:
            rows, cols, sid = (self._pinned("b_" + k, tot, np.int32) for k in ("rows", "cols", "sid"))
            v, p, sig = (self._pinned("b_" + k, tot, np.float64) for k in ("v", "p", "sigma"))
            pp = self._pinned("b_pair", tot, np.float64) if pair else None
        else:
            rows, cols, sid = (np.empty(tot, np.int32) for _ in range(3))
            v, p, sig = (np.empty(tot, np.float64) for _ in range(3))
            pp = np.empty(tot, np.float64) if pair else None
        self._chk(self.lib.mb200_fetch_packed(self.h, tot, _ptr(rows, _i32p), _ptr(cols, _i32p), _ptr(v, _f64p), _ptr(sid, _i32p),
                                              _ptr(p, _f64p), _ptr(sig, _f64p), _ptr(pp, _f64p) if pair else None))
        if sort and tot:
            blk = np.repeat(np.arange(self.nblocks), np.diff(off))
            order = np.lexsort((cols, rows, blk))
            rows, cols, v, p, sid, sig = rows[order], cols[order], v[order], p[order], sid[order], sig[order]
            if pair:
                pp = pp[order]
        out = []
        for b in range(self.nblocks):
            a, z = int(off[b]), int(off[b + 1])
            r = dict(rows=rows[a:z], cols=cols[a:z], v=v[a:z], p=p[a:z], score_id=sid[a:z], sigma=sig[a:z],
                     nz_count=int(nz[b]), n_found=z - a)
            if pair:
                r["pair"] = pp[a:z]
            out.append(r)
        return out

    # ---- device half of the post-processing (mustache.py:774-811) ----
    def select_candidates(self, pt, st, candidate_fraction=-1.0):
        """BH per block, o < pt, sparsity filter and neighbourhood patches on the device (asynchronous)."""
        self._post_args = (float(pt), float(st))
        self._chk(self.lib.mb200_select_candidates(self.h, float(pt), float(st), float(candidate_fraction)))

    def candidates_batch(self, pair=False):
        """One dict per block: rows, cols, q, sigma, cval, keep (sparsity filter passed), o9 / so9 ([m, 9] neighbourhoods of
        the dense o / so matrices), nz_count, n_found; entries sorted row-major.  pair=True (after run_differential): also
        pair9 / vself9 / vother9, the neighbourhoods the differential selection reads.  Only the selected pixels (q < pt)
        cross the bus; a capacity overflow re-runs the selection with room for every record."""
        n_out = C.c_int64(0)
        nul = [None] * 12
        st = self.lib.mb200_fetch_candidates(self.h, 0, *nul, C.byref(n_out))
        if st == -3:
            # MB200_ERR_CAPACITY is either a block with more records than the record capacity (the caller re-runs the
            # batch, blockrun.run_batches) or more selected pixels than the candidate capacity (re-select here, with room
            # for every record)
            msg = self.lib.mb200_last_error(self.h).decode()
            _, found = self.batch_counts()
            if n_out.value > 0 and n_out.value <= int(found.sum()):
                self._chk(self.lib.mb200_select_candidates(self.h, self._post_args[0], self._post_args[1], 1.0))
                st = self.lib.mb200_fetch_candidates(self.h, 0, *nul, C.byref(n_out))
            else:
                raise EngineError(st, msg)
        self._chk(st)
        m = n_out.value
        if m:
            self._chk(self.lib.mb200_enrich_candidates(self.h))       # enrichment filter on the device (flags bits 1-2)
        blk, rows, cols, flg = (self._pinned("c_" + k, m, np.int32) for k in ("blk", "rows", "cols", "flg"))
        q, sg, cv = (self._pinned("c_" + k, m, np.float64) for k in ("q", "sg", "cv"))
        o9, so9 = (self._pinned("c_" + k, 9 * m, np.float64) for k in ("o9", "so9"))
        extra = [self._pinned("c_" + k, 9 * m, np.float64) for k in ("pair9", "vs9", "vo9")] if pair else []
        self._chk(self.lib.mb200_fetch_candidates(self.h, m, _ptr(blk, _i32p), _ptr(rows, _i32p), _ptr(cols, _i32p), _ptr(flg, _i32p),
                                                  _ptr(q, _f64p), _ptr(sg, _f64p), _ptr(cv, _f64p), _ptr(o9, _f64p), _ptr(so9, _f64p),
                                                  *([_ptr(a, _f64p) for a in extra] if pair else [None, None, None]), C.byref(n_out)))
        nz, nf = self.batch_counts()
        order = np.lexsort((cols, rows, blk))
        blk, rows, cols, flg, q, sg, cv = (a[order] for a in (blk, rows, cols, flg, q, sg, cv))
        o9, so9 = o9.reshape(-1, 9)[order], so9.reshape(-1, 9)[order]
        extra = [a.reshape(-1, 9)[order] for a in extra]
        bounds = np.searchsorted(blk, np.arange(self.nblocks + 1))
        out = []
        for b in range(self.nblocks):
            a, z = int(bounds[b]), int(bounds[b + 1])
            d = dict(rows=rows[a:z], cols=cols[a:z], q=q[a:z], sigma=sg[a:z], cval=cv[a:z], keep=(flg[a:z] & 1) != 0,
                     enriched=(flg[a:z] & 2) != 0,
                     o9=o9[a:z], so9=so9[a:z], nz_count=int(nz[b]), n_found=int(nf[b]))
            if pair:
                d.update(pair9=extra[0][a:z], vself9=extra[1][a:z], vother9=extra[2][a:z])
            out.append(d)
        return out

    def q_values(self, block, sort=True):
        """BH-corrected p of every record of a block, in the (sorted) order of records(block)."""
        nz, nf = self.counts(block)
        q = np.empty(nf, np.float64)
        n_out = C.c_int64(0)
        self._chk(self.lib.mb200_fetch_q(self.h, int(block), nf, _ptr(q, _f64p), C.byref(n_out)))
        if sort and nf:
            r = self.records(block, sort=False)
            q = q[np.lexsort((r["cols"], r["rows"]))]
        return q

    def post_ms(self):
        f = C.c_float(0)
        self._chk(self.lib.mb200_last_post_ms(self.h, C.byref(f)))
        return f.value

    def packed_device(self):
        """The packed record arrays as torch CUDA tensors aliasing the engine's buffers (no copy), plus offsets and the
        per-block mask sizes: what the multi-GPU gather hands to NCCL."""
        import torch
        off = self.pack()
        tot = int(off[-1])
        nz, _ = self.batch_counts()
        self.sync()
        ptrs = [C.c_void_p() for _ in range(8)]
        self._chk(self.lib.mb200_packed_device(self.h, *[C.byref(q) for q in ptrs]))

        class _Alias:
            def __init__(self, ptr, n, typestr):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}
        dev = torch.device("cuda", self.device)
        out = {}
        names = ("rows", "cols", "v", "score_id", "scored_index", "p", "sigma", "pair")
        for name, q, ts in zip(names, ptrs, ("<i4", "<i4", "<f8", "<i4", "<i4", "<f8", "<f8", "<f8")):
            if q.value and tot:
                out[name] = torch.as_tensor(_Alias(q.value, tot, ts), device=dev)
            elif name != "pair":
                out[name] = torch.empty(0, dtype=torch.int32 if ts == "<i4" else torch.float64, device=dev)
        out.update(offsets=off, nz_count=nz)
        return out

    def records_device(self, block):
        """Record arrays of a block as torch CUDA tensors aliasing the engine's buffers (no copy): dict of
        rows, cols (int32), v, p (float64), scored_index (int32), each of length n_found; plus nz_count."""
        import torch
        nz, nf = self.counts(block)                       # synchronises the engine's stream
        ptrs = [C.c_void_p() for _ in range(5)]
        cap = C.c_int64(0)
        self._chk(self.lib.mb200_records_device(self.h, int(block), *[C.byref(q) for q in ptrs], C.byref(cap)))

        class _Alias:                                     # __cuda_array_interface__ view of engine-owned device memory
            def __init__(self, ptr, n, typestr):
                self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}
        dev = torch.device("cuda", self.device)
        out = {}
        for name, q, ts in zip(("rows", "cols", "v", "scored_index", "p"), ptrs, ("<i4", "<i4", "<f8", "<i4", "<f8")):
            out[name] = torch.as_tensor(_Alias(q.value, max(nf, 1), ts), device=dev)[:nf]
        out.update(nz_count=nz, n_found=nf)
        return out

    def fits(self, block):
        ns = self.program.n_scored
        loc, sc, sid = np.empty(ns), np.empty(ns), np.empty(ns, np.int32)
        n = C.c_int(0)
        self._chk(self.lib.mb200_fetch_fits(self.h, int(block), _ptr(loc, _f64p), _ptr(sc, _f64p), _ptr(sid, _i32p), ns, C.byref(n)))
        return dict(loc=loc, scale=sc, score_id=sid)

    def timing(self):
        f = [C.c_float(0) for _ in range(6)]
        self._chk(self.lib.mb200_last_timing(self.h, *[C.byref(x) for x in f]))
        return dict(zip(("prep_ms", "kv_ms", "kh_ms", "ks_ms", "fin_ms", "total_ms"), [x.value for x in f]))

    def launches(self):
        n = C.c_int(0)
        self._chk(self.lib.mb200_last_launches(self.h, C.byref(n)))
        return n.value

    def debug_level(self, block, step):
        g = np.zeros((self.n, self.n))
        l = np.zeros((self.n, self.n))
        self._chk(self.lib.mb200_debug_level(self.h, int(block), int(step), C.c_void_p(g.ctypes.data), C.c_void_p(l.ctypes.data)))
        return g, l

    def scale_space_dense(self, tile, dpx, intra=True):
        """One-call drop-in for the scale-space half of mustache(c, ...) on a host tile."""
        n = tile.shape[0]
        cap = max(4096, n * (min(dpx + 1, n - 1) - 3) // 8)
        rows, cols = np.empty(cap, np.int32), np.empty(cap, np.int32)
        v, p, sid = np.empty(cap, np.float64), np.empty(cap, np.float64), np.empty(cap, np.int32)
        nz, nf = C.c_int64(0), C.c_int64(0)
        self._chk(self.lib.mb200_scale_space_dense(self.h, C.c_void_p(tile.ctypes.data), n, tile.strides[0] // 8, int(dpx),
                                                   1 if intra else 0, cap, _ptr(rows, _i32p), _ptr(cols, _i32p),
                                                   _ptr(v, _f64p), _ptr(sid, _i32p), _ptr(p, _f64p), C.byref(nz), C.byref(nf)))
        self.n, self.dpx, self.nblocks = n, int(dpx), 1
        k = nf.value
        order = np.lexsort((cols[:k], rows[:k]))
        sig = self._sigma_lut()[sid[:k][order]]
        return dict(rows=rows[:k][order], cols=cols[:k][order], v=v[:k][order], p=p[:k][order], score_id=sid[:k][order],
                    sigma=sig, nz_count=nz.value, n_found=k)
