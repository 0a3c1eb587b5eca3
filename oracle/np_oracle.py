"""TEST INFRASTRUCTURE ONLY - numpy restatement of the element-wise / scan operators of the gdf
hot path.  Every function cites the reference file:line (under /root/reference/libgdf/src) it follows."""
import numpy as np


def unpack_valid(mask, n):
    """Arrow validity bitmap (LSB first, reference include/gdf/utils.h:10-15) -> bool[n]; None = all valid."""
    if mask is None:
        return np.ones(n, dtype=bool)
    return np.unpackbits(np.asarray(mask, dtype=np.uint8), bitorder="little")[:n].astype(bool)


def pack_valid(bits):
    return np.packbits(np.asarray(bits, dtype=bool), bitorder="little")


# ---- binary ops (reference binaryops.cu:9-31,120-160,290-340,470-500) --------------------------------
def binary_op(name, lhs, rhs, out_init, lvalid=None, rvalid=None):
    """out[i] = lhs[i] OP rhs[i] where both sides are valid, other lanes keep out_init[i]."""
    n = len(lhs)
    with np.errstate(all="ignore"):
        if name == "add": res = lhs + rhs
        elif name == "sub": res = lhs - rhs
        elif name == "mul": res = lhs * rhs
        elif name == "div": res = lhs / rhs
        elif name == "floordiv":
            if np.issubdtype(lhs.dtype, np.integer):   # via double, as the reference does (:143-149)
                res = np.floor(lhs.astype(np.float64) / rhs.astype(np.float64)).astype(lhs.dtype)
            else:
                res = np.floor(lhs / rhs)
        elif name == "gt": res = lhs > rhs
        elif name == "ge": res = lhs >= rhs
        elif name == "lt": res = lhs < rhs
        elif name == "le": res = lhs <= rhs
        elif name == "eq": res = lhs == rhs
        elif name == "ne": res = lhs != rhs
        elif name == "bitwise_and": res = lhs & rhs
        elif name == "bitwise_or": res = lhs | rhs
        elif name == "bitwise_xor": res = lhs ^ rhs
        else: raise ValueError(name)
    res = res.astype(out_init.dtype)
    both = unpack_valid(lvalid, n) & unpack_valid(rvalid, n)
    return np.where(both, res, out_init)


# ---- reductions (reference reductions.cu:26-61,127-190,231-269) ---------------------------------------
def reduce(name, data, valid=None):
    """Reduction over valid rows in the column's own C type (integers wrap)."""
    v = data[unpack_valid(valid, len(data))]
    dt = data.dtype
    with np.errstate(all="ignore"):
        if name == "sum":
            return v.sum(dtype=dt) if len(v) else dt.type(0)
        if name == "product":
            return v.prod(dtype=dt) if len(v) else dt.type(1)
        if name == "sum_squared":
            return (v * v).sum(dtype=dt) if len(v) else dt.type(0)
        if name == "min":
            ident = np.finfo(dt).max if np.issubdtype(dt, np.floating) else np.iinfo(dt).max
            return v.min() if len(v) else dt.type(ident)
        if name == "max":
            ident = np.finfo(dt).min if np.issubdtype(dt, np.floating) else np.iinfo(dt).min
            return v.max() if len(v) else dt.type(ident)
    raise ValueError(name)


# ---- comparisons -> int8 stencil (reference filterops.cu:17-75,97-157) ------------------------------------
# NOTE: the reference's LESS_THAN / LESS_THAN_OR_EQUALS functors compute x > y / x >= y
# (filterops.cu:58-75).  This oracle restates the *documented* operators; DESIGN.md lists the
# divergence and the parity tests against oracle/_ref cover ==, !=, >, >= only.
_CMP = {0: np.equal, 1: np.not_equal, 2: np.less, 3: np.less_equal, 4: np.greater, 5: np.greater_equal}


def comparison(lhs, rhs, op):
    """lhs: column; rhs: column or numpy scalar.  Mixed types compare under C's usual arithmetic
    conversions, which numpy reproduces for these six dtypes except int64-vs-float32 (C: float,
    numpy: float64) - handled explicitly."""
    a, b = np.asarray(lhs), np.asarray(rhs)
    pair = {a.dtype, b.dtype}
    if np.dtype(np.float32) in pair and not (np.dtype(np.float64) in pair):
        a, b = a.astype(np.float32), b.astype(np.float32)
    return _CMP[int(op)](a, b).astype(np.int8)


# ---- gdf_filter (reference sqls_rtti_comp.hpp:200-213,343-370) ----------------------------------------
def filter_rows(cols, vals):
    keep = np.ones(len(cols[0]), dtype=bool)
    for c, v in zip(cols, vals):
        keep &= ~(c != c.dtype.type(v))
    return np.nonzero(keep)[0].astype(np.uint64)


# ---- gpu_apply_stencil (reference streamcompactionops.cu:89-107,148-158,208-339) ----------------------------
def apply_stencil(data, stencil, stencil_valid):
    """Rows kept: stencil byte != 0 AND the stencil validity bit read MSB-first inside each byte
    (bit 7-(i%8): modulus_bit_width with its never-initialised n_bytes, :89-107).  Returns
    (compacted data, output mask bytes) where the mask is what the reference produces when the stencil
    mask is all-valid: ceil(n/8) bytes of ones, a ragged last byte packed MSB-first (:185-198)."""
    n = len(data)
    sv = np.asarray(stencil_valid, dtype=np.uint8)
    i = np.arange(n)
    bit = (sv[i >> 3] >> (7 - (i & 7))) & 1
    keep = (np.asarray(stencil) != 0) & (bit == 1)
    nbytes = (n + 7) // 8
    mask = np.full(nbytes, 0xFF, dtype=np.uint8)
    if n % 8:
        mask[-1] = (0xFF << (8 - n % 8)) & 0xFF
    return data[keep], mask
