"""Host <-> native field marshaling (numpy), i.e. what QUDA's copy_gauge / copy_color_spinor / copy_clover do
when a host field is loaded: QDP-order gauge -> FloatN planes with 18/12/8-parameter packing and the
neighbour's backward links in the pad; SPACE_SPIN_COLOR DeGrand-Rossi spinor <-> FloatN planes in the UKQCD
basis (with block-float int16 for half); packed 72-real clover -> native (optionally compressed) order.

These routines run on the host and exist to feed the engine from oracle-order data in tests, smoke() and
bench.py (device-side marshaling kernels come next).  The layouts are the reference's:
  include/color_spinor_field_order.h:1191-1300, include/kernels/copy_color_spinor.cuh:4-89 (basis rotation),
  include/gauge_field_order.h:1516-1588 (FloatNOrder), :964-1005/:1071-1141/:1267-1310 (Pack for 18/12/8),
  include/clover_field_order.h:79-172 (compression), :587-760 (FloatNOrder), :854-866 (factor 1/2).
"""
import numpy as np

DOUBLE, SINGLE, HALF = 8, 4, 2
_FIXED_MAX = np.float32(32767.0)
_FIXED_INV_MAX = np.float32(3.0518509476e-5)

_S1 = [1, 2, 3, 0]
_S2 = [3, 0, 1, 2]
_K1_NONREL = np.array([+1, -1, -1, -1]) / np.sqrt(2.0)
_K2_NONREL = np.array([+1, -1, +1, +1]) / np.sqrt(2.0)
_K1_REL = np.array([-1, +1, +1, +1]) / np.sqrt(2.0)
_K2_REL = np.array([-1, +1, -1, -1]) / np.sqrt(2.0)


def real_dtype(prec):
    return {DOUBLE: np.float64, SINGLE: np.float32, HALF: np.float32}[prec]


def store_dtype(prec):
    return {DOUBLE: np.float64, SINGLE: np.float32, HALF: np.int16}[prec]


def spinor_N(prec):
    return {DOUBLE: 2, SINGLE: 4, HALF: 8}[prec]


def gauge_N(prec, recon):
    if prec == DOUBLE:
        return 2
    if prec == SINGLE:
        return 2 if recon == 18 else 4
    return {18: 2, 12: 4, 8: 8}[recon]


def volume_cb(X):
    return int(X[0]) * int(X[1]) * int(X[2]) * int(X[3]) // 2


# ---------------------------------------------------------------------------------------------- index maps
def cb_coords(X, parity):
    """coords[x_cb] = (x, y, z, t) for every checkerboard site of `parity` (include/index_helper.cuh:284-302)."""
    Vh = volume_cb(X)
    cb = np.arange(Vh)
    za = cb // (X[0] // 2)
    zb = za // X[1]
    y = za - zb * X[1]
    t = zb // X[2]
    z = zb - t * X[2]
    x = 2 * cb + ((y + z + t + parity) & 1) - za * X[0]
    return np.stack([x, y, z, t], axis=1)


def cb_index(coords, X):
    c = np.asarray(coords)
    return (((c[..., 3] * X[2] + c[..., 2]) * X[1] + c[..., 1]) * X[0] + c[..., 0]) >> 1


def face_sites(X, d, xd, parity):
    """x_cb of the face sites (x[d] == xd, given parity), ordered by face index (index_helper.cuh:445-500)."""
    c = cb_coords(X, parity)
    sel = np.nonzero(c[:, d] == xd)[0]
    cs = c[sel]
    o = [e for e in range(4) if e != d]
    fidx = ((cs[:, o[2]] * X[o[1]] + cs[:, o[1]]) * X[o[0]] + cs[:, o[0]]) >> 1
    out = np.empty(len(sel), dtype=np.int64)
    out[fidx] = sel
    return out


# ---------------------------------------------------------------------------------------------- spinors
def rotate_basis(host, to_nonrel=True):
    """DeGrand-Rossi <-> UKQCD on arrays [..., 4, 3, 2] (copy_color_spinor.cuh:51-89)."""
    K1, K2 = (_K1_NONREL, _K2_NONREL) if to_nonrel else (_K1_REL, _K2_REL)
    h = np.asarray(host, dtype=np.float64)
    out = np.empty_like(h)
    for s in range(4):
        out[..., s, :, :] = K1[s] * h[..., _S1[s], :, :] + K2[s] * h[..., _S2[s], :, :]
    return out


def spinor_bytes(X, prec):
    Vh = volume_cb(X)
    return Vh * 24 * prec + (Vh * 4 if prec == HALF else 0)


def spinor_to_native(host, prec, rotate=True):
    """host [Vh,4,3,2] (DeGrand-Rossi) -> uint8 buffer of one native parity block."""
    Vh = host.shape[0]
    v = rotate_basis(host, True) if rotate else np.asarray(host, dtype=np.float64)
    flat = v.reshape(Vh, 24)
    N = spinor_N(prec)
    if prec == HALF:
        f = flat.astype(np.float32)
        mx = np.abs(f).max(axis=1)
        norm = (mx * _FIXED_INV_MAX).astype(np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            sinv = (_FIXED_MAX / mx).astype(np.float32)
        q = np.rint(f * sinv[:, None]).astype(np.int16)
        planes = np.ascontiguousarray(q.reshape(Vh, 24 // N, N).transpose(1, 0, 2))
        return np.concatenate([planes.view(np.uint8).ravel(), norm.view(np.uint8).ravel()])
    f = flat.astype(real_dtype(prec))
    planes = np.ascontiguousarray(f.reshape(Vh, 24 // N, N).transpose(1, 0, 2))
    return planes.view(np.uint8).ravel().copy()


def spinor_from_native(buf, Vh, prec, rotate=True):
    """inverse of spinor_to_native -> float64 [Vh,4,3,2] (DeGrand-Rossi if rotate)."""
    N = spinor_N(prec)
    raw = np.frombuffer(np.ascontiguousarray(buf).tobytes(), dtype=np.uint8)
    if prec == HALF:
        q = raw[: Vh * 48].view(np.int16).reshape(24 // N, Vh, N).transpose(1, 0, 2).reshape(Vh, 24)
        norm = raw[Vh * 48: Vh * 52].view(np.float32)
        flat = q.astype(np.float32) * norm[:, None]
    else:
        flat = raw[: Vh * 24 * prec].view(real_dtype(prec)).reshape(24 // N, Vh, N).transpose(1, 0, 2).reshape(Vh, 24)
    v = flat.astype(np.float64).reshape(Vh, 4, 3, 2)
    return rotate_basis(v, False) if rotate else v


# ---------------------------------------------------------------------------------------------- gauge
def _pack_links(u, recon):
    """u: [..., 3, 3, 2] float64 -> [..., recon] packed reals (gauge_field_order.h Pack())."""
    flat = u.reshape(u.shape[:-3] + (18,))
    if recon == 18:
        return flat.copy()
    if recon == 12:
        return flat[..., :12].copy()
    out = np.empty(u.shape[:-3] + (8,), dtype=np.float64)
    out[..., 0] = np.arctan2(u[..., 1, 0, 1], u[..., 1, 0, 0]) / np.pi
    out[..., 1] = np.arctan2(-u[..., 2, 0, 1], -u[..., 2, 0, 0]) / np.pi
    out[..., 2:4] = u[..., 1, 1, :]
    out[..., 4:6] = u[..., 1, 2, :]
    out[..., 6:8] = u[..., 0, 0, :]
    return out


def gauge_pad(X):
    """pad >= largest face (tests/utils/set_params.cpp:66-75), in checkerboard sites."""
    V = 2 * volume_cb(X)
    return max(V // X[d] for d in range(4)) // 2


def gauge_boundary_links(host, X, d):
    """Links U_d on the x[d] = X[d]-1 slice, [2][face_cb][3][3][2] in face-index order: what the FORWARD neighbour
    needs in its pad (its backward-hop links across the boundary)."""
    Vh = volume_cb(X)
    g = np.asarray(host).reshape(4, 2, Vh, 3, 3, 2)
    return np.stack([g[d, p][face_sites(X, d, X[d] - 1, p)] for p in range(2)])


def gauge_to_native(host, X, prec, recon, link_max=None, ghost_from=None, ghost_faces=None):
    """host QDP-order gauge [4][V][3][3][2] (parity-major sites) -> native uint8 buffer.

    Layout: [parity][dir*M + i][stride][N], stride = Vh + pad.  The pad of direction d holds, at
    x_cb = Vh + face_idx, the backward neighbour's links U_d on its x[d] = X[d]-1 slice (lib/gauge_field.cpp:453-575);
    `ghost_from[d]` is the host gauge field of that neighbour (default: this field, i.e. periodic self-neighbour);
    alternatively `ghost_faces[d]` gives just those links ([2][face_cb][3][3][2], see gauge_boundary_links).
    Returns (buffer, meta) with meta = dict(stride, parity_stride_bytes, link_max).
    """
    Vh = volume_cb(X)
    pad = gauge_pad(X)
    stride = Vh + pad
    N = gauge_N(prec, recon)
    M = recon // N
    g = np.asarray(host, dtype=np.float64).reshape(4, 2, Vh, 3, 3, 2)
    if link_max is None:
        link_max = float(np.abs(g).max())
    packed = np.zeros((2, 4, stride, recon), dtype=np.float64)
    for mu in range(4):
        src = g if ghost_from is None or ghost_from[mu] is None else \
            np.asarray(ghost_from[mu], dtype=np.float64).reshape(4, 2, Vh, 3, 3, 2)
        for p in range(2):
            packed[p, mu, :Vh] = _pack_links(g[mu, p], recon)
            if ghost_faces is not None and ghost_faces[mu] is not None:
                fl = np.asarray(ghost_faces[mu], dtype=np.float64)[p]
                packed[p, mu, Vh:Vh + len(fl)] = _pack_links(fl, recon)
            else:
                fs = face_sites(X, mu, X[mu] - 1, p)
                packed[p, mu, Vh:Vh + len(fs)] = _pack_links(src[mu, p][fs], recon)
    if prec == HALF:
        scaled = packed / link_max if recon == 18 else packed
        q = np.rint(scaled.astype(np.float32) * _FIXED_MAX).astype(np.int16)
        arr = q
    else:
        arr = packed.astype(real_dtype(prec))
    # [p][mu][stride][M][N] -> [p][mu][M][stride][N]
    arr = np.ascontiguousarray(arr.reshape(2, 4, stride, M, N).transpose(0, 1, 3, 2, 4))
    buf = arr.view(np.uint8).ravel().copy()
    meta = dict(stride=stride, parity_stride_bytes=buf.size // 2, link_max=link_max, pad=pad)
    return buf, meta


# ---------------------------------------------------------------------------------------------- clover
def clover_to_native(host, X, prec, compressed=True, diagonal=None):
    """host packed clover [V][2][36] -> native uint8 buffer [parity][planes][Vh][N] holding A/2.

    Returns (buffer, meta) with meta = dict(parity_stride_bytes, diagonal, max_element, compressed)."""
    Vh = volume_cb(X)
    c = 0.5 * np.asarray(host, dtype=np.float64).reshape(2, Vh, 2, 36)
    if compressed:
        if diagonal is None:
            # the compressed format needs diag[i] + diag[i+3] == 2*diagonal; take it from the data
            diagonal = float(np.mean(0.5 * (c[..., 0:3] + c[..., 3:6])))
        st = np.empty((2, Vh, 2, 28), dtype=np.float64)
        st[..., 0:3] = c[..., 0:3] - diagonal
        st[..., 3] = 0.0
        st[..., 4:28] = c[..., 6:30]
    else:
        diagonal = 0.0 if diagonal is None else diagonal
        st = c
    CB = st.shape[-1]
    N = spinor_N(prec)
    flat = st.reshape(2, Vh, 2 * CB)
    nplanes = 2 * CB // N
    max_element = 2.0 * float(np.abs(flat).max())
    if prec == HALF:
        nrm = np.float32(max_element / (2.0 * 32767.0))
        q = np.rint(flat.astype(np.float32) * (np.float32(1.0) / nrm)).astype(np.int16)
        arr = q
    else:
        arr = flat.astype(real_dtype(prec))
    arr = np.ascontiguousarray(arr.reshape(2, Vh, nplanes, N).transpose(0, 2, 1, 3))
    buf = arr.view(np.uint8).ravel().copy()
    meta = dict(parity_stride_bytes=buf.size // 2, diagonal=diagonal, max_element=max_element,
                compressed=int(compressed))
    return buf, meta


def clover_is_compressible(host, tol=1e-12):
    """True if every chiral block satisfies the symmetry the 28-real format assumes (clover_field_order.h:141-158)."""
    c = np.asarray(host, dtype=np.float64).reshape(-1, 36)
    d = 0.5 * (c[:, 0:3] + c[:, 3:6])
    ok = np.abs(d - d.mean()).max() < tol
    ok &= np.abs(c[:, 30:34] + c[:, 6:10]).max() < tol
    ok &= np.abs(c[:, 34:36] + c[:, 16:18]).max() < tol
    return bool(ok)


# ---------------------------------------------------------------------------------------------- ghosts
def ghost_parity_bytes(X, prec, d):
    face_cb = volume_cb(X) * 2 // X[d] // 2
    return face_cb * (12 * prec + (4 if prec == HALF else 0))


# ---------------------------------------------------------------------------------------------- device-side packing
def gauge_to_native_torch(u, X, prec, recon, ghost_faces=None):
    """Same result as gauge_to_native (fp64 / fp32 / half) but computed with torch ops on the device holding `u`
    ([4][V][3][3][2], parity-major sites).  Used to set up the 32^4 benchmark fields in seconds."""
    import torch
    Vh = volume_cb(X)
    pad = gauge_pad(X)
    stride = Vh + pad
    N = gauge_N(prec, recon)
    M = recon // N
    g = u.reshape(4, 2, Vh, 3, 3, 2).to(torch.float64)
    link_max = float(g.abs().max().item())

    def pack(w):
        flat = w.reshape(w.shape[:-3] + (18,))
        if recon == 18:
            return flat
        if recon == 12:
            return flat[..., :12]
        o = torch.empty(w.shape[:-3] + (8,), dtype=torch.float64, device=w.device)
        o[..., 0] = torch.atan2(w[..., 1, 0, 1], w[..., 1, 0, 0]) / np.pi
        o[..., 1] = torch.atan2(-w[..., 2, 0, 1], -w[..., 2, 0, 0]) / np.pi
        o[..., 2:4] = w[..., 1, 1, :]
        o[..., 4:6] = w[..., 1, 2, :]
        o[..., 6:8] = w[..., 0, 0, :]
        return o

    packed = torch.zeros((2, 4, stride, recon), dtype=torch.float64, device=u.device)
    for mu in range(4):
        for p in range(2):
            packed[p, mu, :Vh] = pack(g[mu, p])
            if ghost_faces is not None and ghost_faces[mu] is not None:
                fl = ghost_faces[mu][p].to(torch.float64)
            else:
                fs = torch.from_numpy(face_sites(X, mu, X[mu] - 1, p)).to(u.device)
                fl = g[mu, p][fs]
            packed[p, mu, Vh:Vh + fl.shape[0]] = pack(fl)
    if prec == HALF:
        scaled = packed / link_max if recon == 18 else packed
        arr = torch.round(scaled.to(torch.float32) * float(_FIXED_MAX)).to(torch.int16)
    else:
        arr = packed.to(torch.float64 if prec == DOUBLE else torch.float32)
    arr = arr.reshape(2, 4, stride, M, N).permute(0, 1, 3, 2, 4).contiguous()
    buf = arr.view(torch.uint8).reshape(-1)
    meta = dict(stride=stride, parity_stride_bytes=buf.numel() // 2, link_max=link_max, pad=pad)
    return buf, meta
