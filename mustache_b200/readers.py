"""Contact-map readers (host side).  Only the text reader is implemented natively; `.hic` / `.cool` / `.mcool`
inputs need the same third-party packages the reference needs (hicstraw, cooler) and are imported lazily, so the
package imports cleanly on images that lack them (this one does).

Restates mustache.py:199-297 (`get_sep`, `read_bias`, `read_pd`).  Out of scope for kernels.
"""
import os

import numpy as np


def strip_chr(name):
    return str(name).replace("chr", "")


def same_chromosome(a, b):
    """mustache.py:191-196."""
    return strip_chr(a) == strip_chr(b)


def guess_separator(path):
    """mustache.py:199-215: decided from the first line only."""
    with open(path) as fh:
        for line in fh:
            if "\t" in line:
                return "\t"
            if " " in line.strip():
                return " "
            if "," in line:
                return ","
            if len(line.split(" ")) == 1:
                return " "
            break
    raise FileNotFoundError(path)


def read_bias(path, chromosome, res):
    """mustache.py:218-251.  Returns {bin: bias}; bias < 0.2 or NaN becomes +inf (the contact is then dropped)."""
    if not path:
        return False
    table = {}
    sep = guess_separator(path)
    with open(path) as fh:
        for pos, line in enumerate(fh):
            parts = line.strip().split(sep)
            if len(parts) == 3:
                if not same_chromosome(parts[0], chromosome):
                    continue
                key, val = float(parts[1]) // res, float(parts[2])
            elif len(parts) == 1:
                key, val = pos, float(parts[0])
            else:
                continue
            table[key] = np.inf if (np.isnan(val) or val < 0.2) else val
    return table


def _bias_factors(table, bins):
    uniq, inv = np.unique(np.asarray(bins), return_inverse=True)
    vals = np.array([table.get(b, 1) for b in uniq.tolist()], dtype=np.float64)
    return vals[inv]


def _chromosome_rows(column, chromosome):
    """np.vectorize(is_chr)(column, chromosome) of mustache.py:259, 263, evaluated once per distinct name instead of once
    per row (a chromosome column holds a handful of names; the per-row Python call was 80 % of the reader's time)."""
    names = [s for s in column.unique().tolist() if same_chromosome(s, chromosome)]
    return column.isin(names).to_numpy(dtype=bool)


def parse_contacts_native(path, chromosome):
    """The parse half of read_pd() through the C ABI (mb200_contacts_open: multi-threaded native parser).  Returns
    (pos1, pos2, value) for the rows of `chromosome`, the value column in the dtype pandas would infer, or None when the
    file holds anything the strict native parser leaves to pandas."""
    import ctypes as C
    from .engine import load_library
    lib = load_library()
    h, n, ncols, is_int = C.c_void_p(), C.c_int64(0), C.c_int(0), C.c_int(0)
    st = lib.mb200_contacts_open(str(path).encode(), str(chromosome).encode(), int(os.environ.get("MUSTACHE_PARSE_THREADS", "0")),
                                 C.byref(h), C.byref(n), C.byref(ncols), C.byref(is_int))
    if st != 0:
        return None
    try:
        a, b, val = np.empty(n.value, np.int64), np.empty(n.value, np.int64), np.empty(n.value, np.float64)
        lib.mb200_contacts_read(h, a.ctypes.data_as(C.POINTER(C.c_int64)), b.ctypes.data_as(C.POINTER(C.c_int64)),
                                val.ctypes.data_as(C.POINTER(C.c_double)))
    finally:
        lib.mb200_contacts_close(h)
    return a, b, (val.astype(np.int64) if is_int.value else val), ncols.value


def _parse_contacts_pandas(path, chromosome):
    """pd.read_csv + dropna + chromosome filter exactly as mustache.py:255-263 (the fallback of the native parser)."""
    import pandas as pd
    sep = guess_separator(path)
    df = pd.read_csv(path, sep=sep, header=None)
    df = df.dropna()
    if df.shape[1] == 5:
        df = df[_chromosome_rows(df[0], chromosome)]
        if df.shape[0] == 0:
            return None
        df = df[_chromosome_rows(df[2], chromosome)]
        return df[1].to_numpy(), df[3].to_numpy(), df[4].to_numpy(), 5
    if df.shape[1] == 3:
        return df[0].to_numpy(), df[1].to_numpy(), df[2].to_numpy(), 3
    raise ValueError("expected a 3- or 5-column contact file, got %d columns" % df.shape[1])


def read_text(path, distance_in_bp, bias_path, chromosome, res):
    """mustache.py:254-297 (`read_pd`): 5-column (chr pos chr pos count) or 3-column (pos pos count) text.

    Returns upper-triangular COO (x, y, value) in bin units: |pos1-pos2| <= (distance/res + 1)*res, divided by both
    biases (in that order), non-positive values dropped.  The value column keeps the dtype pandas inferred, as in the
    reference: integer counts without a bias file stay int64, and normalize_sparse then writes its z-scores into that
    integer array, truncating them (mustache.py:668, 683) -- part of the reference's observable behaviour.
    """
    limit = (distance_in_bp / res + 1) * res
    got = None
    if os.environ.get("MUSTACHE_READER", "native") != "pandas":
        got = parse_contacts_native(path, chromosome)
    if got is None:
        got = _parse_contacts_pandas(path, chromosome)
        if got is None:
            print("Could't read any interaction for this chromosome!")
            return None
    a, b, val, ncols = got
    if ncols == 5 and len(a) == 0:
        print("Could't read any interaction for this chromosome!")
        return None
    near = np.abs(a - b) <= limit
    a, b, val = a[near] // res, b[near] // res, val[near]
    table = read_bias(bias_path, chromosome, res)
    if table:
        val = np.divide(val, _bias_factors(table, a))
        val = np.divide(val, _bias_factors(table, b))
    pos = val > 0
    a, b, val = a[pos], b[pos], val[pos]
    return np.minimum(a, b), np.maximum(a, b), np.array(val)


def read_hic(path, norm_method, chrom_size, distance_in_bp, chr1, chr2, res):  # pragma: no cover - dependency absent
    """mustache.py:300-396 needs `hicstraw`; not installed in this image."""
    try:
        import hicstraw  # noqa: F401
    except ImportError as e:
        raise ImportError(".hic input needs the `hicstraw` package (as the reference does)") from e
    raise NotImplementedError(".hic reading is outside the accelerated path; see DESIGN.md (out of scope)")


def read_cool(path, distance_in_bp, chr1, chr2, norm_method, res=None):  # pragma: no cover - dependency absent
    """mustache.py:399-592 needs `cooler`; not installed in this image."""
    try:
        import cooler  # noqa: F401
    except ImportError as e:
        raise ImportError(".cool/.mcool input needs the `cooler` package (as the reference does)") from e
    raise NotImplementedError(".cool reading is outside the accelerated path; see DESIGN.md (out of scope)")
