"""checkasm-style coefficient generator for itxfm_add.

Follows the recipe of the reference's tests/checkasm/itx.c:183-240 (`ftx`): a
random residual in [-bitdepth_max, bitdepth_max] is pushed through a floating
point forward transform of the matching 1-D kinds, scaled per block area
(itx.c:74-84) and rounded, so that the coefficients have the range real
streams produce.  `subsh` variants keep only a top-left sub-block non-zero
(itx.c:131-181) to exercise dc-only and partial paths.
"""
import numpy as np

from ..lib import TX_DIMS

# txtp -> (row kind, col kind); kinds: 0 dct, 1 adst, 2 flipadst, 3 identity, 4 wht
# (tests/checkasm/itx.c:46-64; src/levels.rs:63-82)
TXTP_KINDS = [(0, 0), (0, 1), (1, 0), (1, 1), (0, 2), (2, 0), (2, 2), (2, 1), (1, 2), (3, 3), (3, 0), (0, 3),
              (3, 1), (1, 3), (3, 2), (2, 3), (4, 4)]

_SCALE = [4.0, 4.0 * 0.5 ** 0.5, 2.0, 2.0 * 0.5 ** 0.5, 1.0, 0.5 * 0.5 ** 0.5, 0.25, 0.125 * 0.5 ** 0.5, 0.0625]


def _fwd_matrix(kind, n):
    j = np.arange(n)[None, :]
    i = np.arange(n)[:, None]
    if kind == 0:
        m = np.cos(np.pi * (2 * j + 1) * i / (2.0 * n))
        m[0, :] *= 0.5 ** 0.5
        return m
    if kind in (1, 2):
        if n == 4:
            return np.sin(np.pi * (j + 1) * (2 * i + 1) / 9.0)
        return np.sin(np.pi * (2 * j + 1) * (2 * i + 1) / (4.0 * n))
    if kind == 4:
        # forward WHT used by checkasm (itx.c:117-129), as a matrix
        m = np.zeros((4, 4))
        for k in range(4):
            e = np.zeros(4); e[k] = 1
            t0 = e[0] + e[1]; t3 = e[3] - e[2]; t4 = (t0 - t3) * 0.5
            t1 = t4 - e[1]; t2 = t4 - e[2]
            m[:, k] = [t0 - t2, t2, t3 + t1, t1]
        return m
    return np.eye(n)


def valid_txtps(tx):
    w, h = TX_DIMS[tx]
    m = max(w, h)
    if m == 64:
        return [0]
    if m == 32:
        return [0, 9]
    if (w, h) == (16, 16):
        return list(range(12))
    return list(range(16)) + ([16] if tx == 0 else [])


def gen_coefs(rng, tx, txtp, bitdepth_max, n=1, variant="full"):
    """Returns (coef[n, sw*sh] column-major (x*sh + y), eob[n]) as int64/int32 arrays.

    variant: "dc" (only coeff 0, eob 0), "sub" (top-left 8x8 or 4x4 non-zero),
             "full", "extreme" (uniform over the legal coefficient range +-(128<<bpc),
             src/recon.rs:1417, to exercise the intermediate clips)."""
    w, h = TX_DIMS[tx]
    sw, sh = min(w, 32), min(h, 32)
    rk, ck = TXTP_KINDS[txtp]
    if variant == "extreme":
        lim = 128 * (bitdepth_max + 1)
        c = rng.integers(-lim, lim, size=(n, sw, sh))
    else:
        res = rng.integers(-bitdepth_max, bitdepth_max + 1, size=(n, h, w)).astype(np.float64)
        mr, mc = _fwd_matrix(rk, w), _fwd_matrix(ck, h)
        scale = _SCALE[int(np.log2(w * h)) - 4]
        t = np.einsum("ij,nyj->nyi", mr, res) * scale      # rows
        o = np.einsum("ij,njx->nix", mc, t)                # cols -> [n, h, w]
        c = np.floor(o[:, :sh, :sw] + 0.5).astype(np.int64).transpose(0, 2, 1)  # [n, x, y]
        lim = 128 * (bitdepth_max + 1)
        c = np.clip(c, -lim, lim - 1)
    c = np.ascontiguousarray(c)
    if variant == "dc":
        c[:, 1:, :] = 0
        c[:, 0, 1:] = 0
    elif variant == "sub":
        k = 8 if max(sw, sh) > 8 else 2
        c[:, k:, :] = 0
        c[:, :, k:] = 0
    flat = c.reshape(n, sw * sh)
    nz = flat != 0
    last = np.where(nz.any(axis=1), flat.shape[1] - 1 - np.argmax(nz[:, ::-1], axis=1), 0)
    eob = last.astype(np.int32)
    if variant == "dc":
        eob[:] = 0
    else:
        eob = np.maximum(eob, 1)  # any eob >= 1 selects the full path (src/itx.rs:90)
    return flat, eob
