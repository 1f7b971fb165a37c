"""Element-subset bit masks: the wire format of the batched solve (uint32 words, bit e%32 of
word e//32 set = element e exists)."""
import numpy as np

USE_NATIVE = True                      # tests switch this off to compare with the numpy path
_cmpack = None


def _native_packer():
    """The CPython packer (csrc_host/cmpack.c), imported on first use: a fresh clone builds it
    (conmech_b200.build) after this module may already have been imported.  Host-side helper only;
    the numpy path below is equivalent."""
    global _cmpack
    if _cmpack is None and USE_NATIVE:
        try:
            from . import _cmpack as mod
            _cmpack = mod
        except ImportError:
            return None
    return _cmpack if USE_NATIVE else None


def ids_to_mask(ids, nE):
    m = np.zeros((nE + 31) // 32, dtype=np.uint32)
    ids = np.asarray(ids, dtype=np.int64)
    np.bitwise_or.at(m, ids >> 5, (np.uint32(1) << (ids & 31).astype(np.uint32)))
    return m


def pack_masks(list_of_ids, nE, out=None):
    """(B, ceil(nE/32)) uint32 masks for a list of id lists.  An empty list means the full
    structure (reference ``numpy_stiffness.py:230-231``).  Ids out of range raise IndexError.
    A repeated id inside one subset raises ValueError: the reference assembles such an element
    twice (``numpy_stiffness_assembly.py:371-388``, no de-duplication), a bit mask is a set and
    cannot express that, and silently solving a different structure is not an option.

    Vectorised over the whole batch (one flat id array, one scatter, one packbits): the batched
    solve consumes several hundred thousand subsets per second and GPU, a Python loop over the
    subsets does not keep up."""
    B = len(list_of_ids)
    W = (nE + 31) // 32
    if B == 0:
        return np.zeros((0, W), dtype=np.uint32) if out is None else out
    native = _native_packer()
    if native is not None and not isinstance(list_of_ids, np.ndarray):
        # native packer (csrc_host/cmpack.c, built by conmech_b200.build): ~5 ns per id straight from
        # the Python lists; the numpy path below (same result) costs ~70 ns per id on lists
        dst = out if (out is not None and out.dtype == np.uint32 and out.flags.c_contiguous and out.shape == (B, W)) \
            else np.empty((B, W), dtype=np.uint32)
        native.pack(list_of_ids, int(nE), dst)
        if out is not None and dst is not out:
            out[...] = dst
            return out
        return dst
    lens = np.fromiter((len(ids) for ids in list_of_ids), dtype=np.int64, count=B)
    total = int(lens.sum())
    if total:      # (np.asarray per list is ~3x faster than one np.fromiter over the chained lists)
        flat = np.concatenate([np.asarray(ids, dtype=np.int64).ravel() for ids in list_of_ids if len(ids)])
    else:
        flat = np.zeros(0, dtype=np.int64)
    if total and (flat.min() < 0 or flat.max() >= nE):
        raise IndexError('element id not within range!')
    bits = np.zeros((B, W * 32), dtype=np.uint8)
    if total:
        bits.reshape(-1)[np.repeat(np.arange(B, dtype=np.int64) * (W * 32), lens) + flat] = 1
        if int(bits.sum()) != total:
            raise ValueError('duplicate element id in a subset (the reference would double count it)')
    if (lens == 0).any():
        bits[lens == 0, :nE] = 1
    packed = np.packbits(bits.reshape(B, W, 32), axis=2, bitorder='little')      # (B, W, 4) bytes
    packed = np.ascontiguousarray(packed).view('<u4').reshape(B, W)
    if out is not None:
        out[...] = packed
        return out
    return packed


def pack_masks_csr(flat_ids, offsets, nE, out=None):
    """Masks for subsets given as two arrays: subset b = ``flat_ids[offsets[b]:offsets[b+1]]``
    (an empty range = the full structure).  The cheapest host format next to the masks
    themselves: no Python object per id (about 1 ns per id in the native packer)."""
    flat = np.ascontiguousarray(flat_ids, dtype=np.int64)
    off = np.ascontiguousarray(offsets, dtype=np.int64)
    B, W = off.size - 1, (nE + 31) // 32
    if B < 0:
        raise ValueError('offsets must have B+1 entries')
    dst = out if (out is not None and out.dtype == np.uint32 and out.flags.c_contiguous and out.shape == (B, W)) \
        else np.empty((B, W), dtype=np.uint32)
    native = _native_packer()
    if native is not None:
        native.pack_csr(flat, off, int(nE), dst)
    else:
        dst[...] = pack_masks([flat[off[b]:off[b + 1]] for b in range(B)], nE)
    if out is not None and dst is not out:
        out[...] = dst
        return out
    return dst


def publish(flags, k, value=1):
    """``flags[k] = value`` with release ordering (int32 array shared with a native consumer thread)."""
    native = _native_packer()
    if native is not None and hasattr(native, 'publish'):
        native.publish(flags, int(k), int(value))
    else:
        flags[k] = value


def as_mask_matrix(a, mask_words):
    """A 2-D array of 4-byte integers with ``mask_words`` columns is a pre-packed mask matrix
    (uint32, or int32 - what torch tensors hold); returns it as C-contiguous uint32.  A 2-D array of
    4-byte integers with another column count is ambiguous (masks of another model? rows of ids?) and
    raises; anything else (lists, int64 id rows) returns None = "a sequence of id sequences"."""
    if not isinstance(a, np.ndarray) or a.ndim != 2 or a.dtype.kind not in 'iu' or a.dtype.itemsize != 4:
        return None
    if a.shape[1] == mask_words:
        return np.ascontiguousarray(a).view(np.uint32)
    raise ValueError('2-D {} array of shape {} is not a mask matrix of this model ((B, {}) uint32/int32); pass id rows '
                     'as int64 or as lists'.format(a.dtype, a.shape, mask_words))


def mask_to_ids(mask, nE):
    bits = np.unpackbits(np.ascontiguousarray(mask).view(np.uint8), bitorder='little')[:nE]
    return np.flatnonzero(bits)
