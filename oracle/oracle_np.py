"""oracle/oracle_np.py — CPU restatement (numpy) of the reference's algorithm for the hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this; the product never does.

Every function restates one reference function (file:line relative to /root/reference/source/lib) on the CTC branch the
BASELINE configs exercise: one or two maps (single stream, or multiple streams with an absolute or delta-coded second
attribute map), singleMapPixelInterleaving and pointLocalReconstruction (their transferColorWeight colours are flagged
exact only where they do not hinge on nanoflann's order of equidistant neighbours), raw patches in the atlas or in the
auxiliary video, no EOM / PBF (those raise NotImplementedError here — the compiled reference in oracle/_ref covers them); grid geometry smoothing and the
non-grid smoothPointCloud (PCCCodec.cpp:1106-1157: the vendored IndexDist_Sorter orders equal distances by index, so no
tree order is involved); plus the decoder-side ingest: PCCImage::set and the 4:2:0 -> 4:4:4 16-bit conversion of
PCCInternalColorConverter (all eight upsampling filters).
It is pinned against the unmodified reference (oracle/_ref/librabbit_ref.so) by tests/test_oracle_port_cpu.py, stage
by stage and bit for bit, and against the golden fixtures in tests/golden/ — so it is a usable checker on a box that
has neither /root/reference nor oracle/_ref.  Not restated here: PCCPointSet3::transferColors16bitBP (needs the
nanoflann traversal order; covered by oracle/_ref and by the GPU kd-tree tests) and the D2 (point-to-plane) metric.
"""
import numpy as np

SWITCHED = (1, 2, 4, 6, 8)  # orientations that exchange U and V on the canvas (PCCPatch.h:63-70)


# ---------------------------------------------------------------------------------------------------------------
# PCCPatch::patch2Canvas (PccLibCommon/source/PCCPatch.cpp:192-251) / patchBlock2CanvasBlock (:253-308)
# ---------------------------------------------------------------------------------------------------------------
def patch2canvas(orient, u, v, su, sv, x0, y0):
    """patch-local pixel (u, v) -> canvas pixel; su, sv = patch size in pixels, (x0, y0) = patch origin in pixels"""
    if orient == 0:
        x, y = u, v
    elif orient in (1, 8):
        x, y = v, u
    elif orient == 2:
        x, y = sv - 1 - v, u
    elif orient == 3:
        x, y = su - 1 - u, sv - 1 - v
    elif orient == 4:
        x, y = v, su - 1 - u
    elif orient == 5:
        x, y = su - 1 - u, v
    elif orient == 6:
        x, y = sv - 1 - v, su - 1 - u
    elif orient == 7:
        x, y = u, sv - 1 - v
    else:
        raise ValueError(orient)
    return x + x0, y + y0


# ---------------------------------------------------------------------------------------------------------------
# PCCCodec::generateOccupancyMap (PccLibCommon/source/PCCCodec.cpp:1584-1606) + the up-sampling of
# generatePointCloud (:557-570)
# ---------------------------------------------------------------------------------------------------------------
def occupancy_map(video, precision, threshold):
    """video: u8 [H/p][W/p].  The reference thresholds the video sample IN PLACE once per full-resolution pixel
    (:1597-1600), i.e. p*p times: with threshold >= 1 and p > 1 the second visit sees the binarised 0/1 and clears it."""
    v = video.astype(np.int64)
    if precision == 1 or threshold == 0:
        lo = (v > threshold)
    else:
        lo = np.zeros_like(v, bool)
    return np.repeat(np.repeat(lo, precision, axis=0), precision, axis=1).astype(np.uint8)


# ---------------------------------------------------------------------------------------------------------------
# PCCCodec::generateBlockToPatchFromOccupancyMapVideo (:1725-1763)
# ---------------------------------------------------------------------------------------------------------------
def block_to_patch(occ, patches, R=16):
    H, W = occ.shape
    blocks = occ.reshape(H // R, R, W // R, R).any(axis=(1, 3))
    b2p = np.zeros((H // R, W // R), np.uint32)
    for i, p in enumerate(patches):
        cw, ch = (p["size_v0"], p["size_u0"]) if p["orientation"] in SWITCHED else (p["size_u0"], p["size_v0"])
        sub = blocks[p["v0"]:p["v0"] + ch, p["u0"]:p["u0"] + cw]
        b2p[p["v0"]:p["v0"] + ch, p["u0"]:p["u0"] + cw][sub] = i + 1  # the last patch that covers an occupied block wins
    return b2p


# ---------------------------------------------------------------------------------------------------------------
# PCCCodec::identifyBoundaryPoints (:266-325), evaluated for every pixel of the frame at once
# ---------------------------------------------------------------------------------------------------------------
def boundary_map(occ):
    H, W = occ.shape
    o = np.pad(occ != 0, 2, constant_values=True)  # outside the image counts as occupied; borders are explicit

    def all_occupied(r):
        acc = np.ones((H, W), bool)
        for dy in range(-r, r + 1):
            for dx in range(-r, r + 1):
                acc &= o[2 + dy:2 + dy + H, 2 + dx:2 + dx + W]
        return acc
    ys, xs = np.mgrid[0:H, 0:W]
    border1 = (xs == 0) | (ys == 0) | (xs == W - 1) | (ys == H - 1)
    border2 = (xs == 1) | (ys == 1) | (xs == W - 2) | (ys == H - 2)
    return (border1 | ~all_occupied(1) | ~all_occupied(2) | border2).astype(np.uint16)


# ---------------------------------------------------------------------------------------------------------------
# PCCCodec::generatePointCloud (:517-978) + generatePoints (:327-515, default branch :497-513) +
# PCCPatch::generatePoint (PCCPatch.h:177-207) + colorPointCloud (:1308-1449, single-stream branch :1418-1422)
# ---------------------------------------------------------------------------------------------------------------
def reconstruct_frame(params, occ_video, geometry, attribute, patches, plr=None, patch_base=0, raw_patches=None,
                      aux_geometry=None, aux_attribute=None):
    """returns a dict of arrays in PCCPointSet3 layouts, in the reference's emission order; plr / patch_base: the point
    local reconstruction tables of the GOF (rb200_plr layout) and the GOF index of this frame's first patch; raw_patches:
    the frame's raw (missed-point) patches, read from the atlas or — use_aux_separate_video — from the auxiliary frames"""
    P = params
    if P.enhanced_occupancy_map_code or P.pbf_enable or P.enable_size_quantization:
        raise NotImplementedError("oracle_np restates the default CTC branch only (see the module docstring)")
    if P.use_additional_points_patch and (P.single_map_pixel_interleaving or P.point_local_reconstruction):
        raise NotImplementedError("oracle_np: raw patches together with pixel interleaving / PLR")
    if P.single_map_pixel_interleaving or P.point_local_reconstruction:
        return reconstruct_frame_interleaved(params, occ_video, geometry, attribute, patches, plr, patch_base)
    R, M = P.occupancy_resolution, P.map_count_minus1 + 1
    occ = occupancy_map(occ_video, P.occupancy_precision, P.threshold_lossy_om)
    b2p = block_to_patch(occ, patches, R)
    btype = boundary_map(occ) if P.flag_geometry_smoothing else np.zeros(occ.shape, np.uint16)
    order = range(len(patches) - 1, -1, -1) if P.patch_precedence_reverse else range(len(patches))
    pos, col, typ, part, p2p = [], [], [], [], []
    for i in order:
        p = patches[i]
        su0, sv0 = int(p["size_u0"]), int(p["size_v0"])
        if su0 == 0 or sv0 == 0:
            continue
        # emission order: v0, u0 (blocks), then v1, u1 (pixels), then the layers (:648-840)
        v0, u0, v1, u1 = np.meshgrid(np.arange(sv0), np.arange(su0), np.arange(R), np.arange(R), indexing="ij")
        u, v = (u0 * R + u1).ravel(), (v0 * R + v1).ravel()
        x, y = patch2canvas(int(p["orientation"]), u, v, su0 * R, sv0 * R, int(p["u0"]) * R, int(p["v0"]) * R)
        keep = (b2p[y // R, x // R] == i + 1) & (occ[y, x] != 0)  # :650, :668
        u, v, x, y = u[keep], v[keep], x[keep], y[keep]
        n = len(u)
        d0 = geometry[0][y, x].astype(np.int64)

        def normal(depth):  # PCCPatch::generateNormalCoordinate, PCCPatch.h:177-186
            return depth + int(p["d1"]) if p["projection_mode"] == 0 else np.maximum(int(p["d1"]) - depth, 0)
        P0 = np.zeros((n, 3), np.int64)
        P0[:, p["normal_axis"]] = normal(d0)
        P0[:, p["tangent_axis"]] = u * int(p["lod_x"]) + int(p["u1"])
        P0[:, p["bitangent_axis"]] = v * int(p["lod_y"]) + int(p["v1"])
        P0 = P0.astype(np.int16)  # double -> int16 (PCCMath.h:79-84)
        layers = [P0]
        if M > 1:
            g1 = geometry[1][y, x].astype(np.int64)
            P1 = P0.copy()
            if P.absolute_d1:
                P1[:, p["normal_axis"]] = normal(g1).astype(np.int16)  # generatePoint( u, v, frame1 ), :503
            else:
                n0 = P0[:, p["normal_axis"]].astype(np.int64)
                P1[:, p["normal_axis"]] = (n0 + g1 if p["projection_mode"] == 0 else n0 - g1).astype(np.int16)  # :505-509
            layers.append(P1)
        # interleave the layers per pixel; D1 == D0 is dropped when removeDuplicatePoints (:794-795)
        emit = np.ones((n, len(layers)), bool)
        if M > 1 and P.remove_duplicate_points:
            emit[:, 1] = (layers[1] != layers[0]).any(axis=1)
        idx = np.nonzero(emit.ravel())[0]
        pix, lay = idx // len(layers), idx % len(layers)
        allp = np.stack(layers, axis=1).reshape(-1, 3)
        pos.append(allp[idx])
        typ.append(btype[y[pix], x[pix]])
        part.append(np.full(len(idx), i, np.uint32))
        p2p.append(np.stack([x[pix], y[pix], lay], axis=1).astype(np.uint32))
        if P.attribute_count > 0:
            c16 = np.stack([attribute[lay, c, y[pix], x[pix]] for c in range(3)], axis=1).astype(np.uint16)
            if P.multiple_streams and P.relative_t1 and M > 1:
                # colorPointCloud, multiple streams with !absoluteT1List[1] (:1387-1416): T1 = T0 + clamped delta
                bits = 8 if P.attribute_rgb444 else 16
                offset, maxv = 1 << (bits - 1), (1 << bits) - 1
                v0 = np.stack([attribute[0, c, y[pix], x[pix]] for c in range(3)], axis=1).astype(np.int64)
                nv = np.clip(c16.astype(np.int64) - offset, -offset, offset - 1) + v0
                c16 = np.where((lay == 1)[:, None], np.clip(nv, 0, maxv), c16).astype(np.uint16)
            col.append(c16)
    # raw (missed-point) patches (:894-949): 3 n consecutive samples of the patch rectangle as X || Y || Z runs, read from
    # map 0 of the atlas or from the auxiliary geometry video; pixel address = the n first pixels of the rectangle;
    # partition = patches.size(); colours from attribute frame 0 at those addresses (:1418-1422) or, with the auxiliary
    # video, from its attribute frames through 8-bit PCCColor3B values (:1524-1549, :1438)
    if P.use_additional_points_patch and raw_patches is not None:
        aux = bool(P.use_aux_separate_video)
        src = aux_geometry if aux else geometry[0]
        for r in raw_patches:
            nr, su, sv = int(r["num_points"]), int(r["size_u0"]) * R, int(r["size_v0"]) * R
            x0, y0 = int(r["u0"]) * R, int(r["v0"]) * R
            vals = src[y0:y0 + sv, x0:x0 + su].reshape(-1)[:3 * nr].astype(np.int64)
            q = np.stack([vals[:nr] + int(r["u1"]), vals[nr:2 * nr] + int(r["v1"]), vals[2 * nr:3 * nr] + int(r["d1"])], axis=1)
            k = np.arange(nr)
            px, py = x0 + k % su, y0 + k // su
            pos.append(q.astype(np.int16))
            typ.append(np.zeros(nr, np.uint16))
            part.append(np.full(nr, len(patches), np.uint32))
            p2p.append(np.stack([px, py, np.zeros(nr, np.int64)], axis=1).astype(np.uint32))
            if P.attribute_count > 0:
                if aux:
                    col.append((np.stack([aux_attribute[c, py, px] for c in range(3)], axis=1) & 0xFF).astype(np.uint16))
                else:
                    col.append(np.stack([attribute[0, c, py, px] for c in range(3)], axis=1).astype(np.uint16))

    def cat(parts, shape, dt):
        return np.concatenate(parts) if parts else np.zeros(shape, dt)
    n = sum(len(a) for a in pos)
    out = dict(positions=cat(pos, (0, 3), np.int16), boundary_types=cat(typ, (0,), np.uint16).astype(np.uint16),
               partition=cat(part, (0,), np.uint32), point_to_pixel=cat(p2p, (0, 3), np.uint32),
               colors16=cat(col, (0, 3), np.uint16) if P.attribute_count > 0 else np.zeros((n, 3), np.uint16),
               colors=np.zeros((n, 3), np.uint8))
    return out, b2p, occ


# ---------------------------------------------------------------------------------------------------------------
# singleMapPixelInterleaving: generatePoints (:350-471); pointLocalReconstruction: generatePoints (:472-496) with
# getDeltaNeighbors (:238-264); the caller's bookkeeping (:781-835) and the colours
# (colorPointCloud :1367-1374, :1429-1434 -> PCCPointSet3::transferColorWeight, PCCPointSet.cpp:2250-2280)
# ---------------------------------------------------------------------------------------------------------------
def _round_half_away(x):
    return np.where(x >= 0, np.floor(x + 0.5), -np.floor(-x + 0.5))


def _interleaved_pixels(p, i, geo, occ, b2p, x, y, R, st):
    """generatePoints :350-471 for the occupied pixels (x, y) of patch i: normal coordinate of the coded point, of the
    interpolated one, whether it exists, and the fill run (start - 1, length)"""
    H, W = occ.shape
    n, d1, mode = len(x), int(p["d1"]), int(p["projection_mode"])
    own = (geo[y, x] + d1 if mode == 0 else np.maximum(d1 - geo[y, x], 0)).astype(np.int16).astype(np.float64)
    # the four neighbours in the reference's order: left, right, top, bottom (:380-433)
    dn = np.zeros((4, n))
    cnt = np.zeros(n, np.int64)
    mn, mx = own.copy(), own.copy()
    for k, (dx, dy) in enumerate(((-1, 0), (1, 0), (0, -1), (0, 1))):
        nx, ny = x + dx, y + dy
        ok = (nx >= 0) & (ny >= 0) & (nx < W) & (ny < H)
        cx, cy = np.clip(nx, 0, W - 1), np.clip(ny, 0, H - 1)
        ok &= (occ[cy, cx] != 0) & (b2p[cy // R, cx // R] == i + 1)
        val = geo[cy, cx]
        # size_t arithmetic: d1 - value wraps to a huge value for value > d1 (:385-389)
        if mode == 0:
            dk = (val + d1).astype(np.float64)
        else:
            diff = d1 - val
            dk = np.where(diff >= 0, diff.astype(np.float64), diff.astype(np.float64) + 18446744073709551616.0)
        dn[k] = np.where(ok, dk, 0.0)
        cnt += ok
        mn = np.where(ok, np.minimum(mn, dk), mn)
        mx = np.where(ok, np.maximum(mx, dk), mx)
    two = cnt > 0  # :434
    odd = ((x + y) & 1) == 1
    with np.errstate(divide="ignore", invalid="ignore"):
        avg = (((dn[0] + dn[1]) + dn[2]) + dn[3]) / cnt
    if mode == 0:
        other = np.where(odd, np.minimum(np.maximum(mn, own - st), own), np.maximum(np.minimum(avg, own + st), own))
    else:
        other = np.where(odd, np.maximum(np.minimum(mx, own + st), own), np.minimum(np.maximum(avg, own - st), own))
    other = np.where(two, _round_half_away(np.where(two, other, 0.0)), 0.0).astype(np.int64).astype(np.int16).astype(np.int64)
    ownI = own.astype(np.int64)
    lo, hi = np.minimum(ownI, other), np.maximum(ownI, other)
    nfill = np.where(two, np.maximum(hi - lo - 1, 0), 0)
    return ownI, other, two, lo, nfill


def _plr_pixels(p, plr, gi, geo, u, v, x, y, R):
    """generatePoints :472-496 + getDeltaNeighbors :238-264 for the occupied pixels of GOF patch gi"""
    H, W = geo.shape
    d1, mode, su0 = int(p["d1"]), int(p["projection_mode"]), int(p["size_u0"])
    m = plr["modes"][plr["block_mode"][plr["block_offset"][gi] + (v // R) * su0 + (u // R)]].astype(np.int64)
    interp, filling, min_d1, nb = m[:, 0] != 0, m[:, 1] != 0, m[:, 2], m[:, 3]

    def normal(depth):
        return depth + d1 if mode == 0 else np.maximum(d1 - depth, 0)
    flat = geo.ravel()
    d_org = normal(geo[y, x])
    delta = np.zeros(len(x), np.int64)
    r = int(nb.max(initial=0))
    for dx in range(-r, r + 1):
        for dy in range(-r, r + 1):
            xx, yy = x + dx, y + dy
            # inclusive upper bounds (:249-252): x == W is the first sample of the next row; past the plane: skipped
            idx = yy * W + xx
            ok = interp & (abs(dx) <= nb) & (abs(dy) <= nb) & (xx >= 0) & (yy >= 0) & (xx <= W) & (yy <= H) & (idx < H * W)
            d = normal(flat[np.clip(idx, 0, H * W - 1)]) - d_org
            if mode == 0:
                delta = np.where(ok & (d <= 4) & (d > delta), d, delta)  # g_neighborThreshold
            else:
                delta = np.where(ok & (d >= -4) & (d < delta), d, delta)
    delta = np.where(delta != 0, delta + (-1 if mode == 0 else 1), 0)  # :261
    delta = np.maximum(delta, min_d1) if mode == 0 else np.minimum(delta, -min_d1)  # :478-482
    ownI = d_org.astype(np.int16).astype(np.int64)
    two = delta != 0
    other = (ownI + delta).astype(np.int16).astype(np.int64)
    lo, hi = np.minimum(ownI, other), np.maximum(ownI, other)
    # size_t xmin / xmax (:487-488): a negative minimum wraps to a huge start, except -1, whose successor is 0
    nfill = np.where(two & filling & (lo >= -1), np.maximum(hi - lo - 1, 0), 0)
    return ownI, other, two, lo, nfill


def reconstruct_frame_interleaved(params, occ_video, geometry, attribute, patches, plr=None, patch_base=0):
    """as reconstruct_frame; additionally returns, under "colors16_exact", which colours do not depend on nanoflann's
    traversal order (the 5 nearest coded points are strictly ordered by distance and separated from the 6th)"""
    from scipy.spatial import cKDTree
    P = params
    assert P.map_count_minus1 == 0
    use_plr = bool(P.point_local_reconstruction) and not P.single_map_pixel_interleaving  # :350 is tested before :472
    assert not use_plr or plr is not None
    R, st = P.occupancy_resolution, float(P.surface_thickness)
    occ = occupancy_map(occ_video, P.occupancy_precision, P.threshold_lossy_om)
    H, W = occ.shape
    b2p = block_to_patch(occ, patches, R)
    btype = boundary_map(occ) if P.flag_geometry_smoothing else np.zeros(occ.shape, np.uint16)
    order = range(len(patches) - 1, -1, -1) if P.patch_precedence_reverse else range(len(patches))
    geo = geometry[0].astype(np.int64)
    pos, col, typ, part, p2p, src = [], [], [], [], [], []
    for i in order:
        p = patches[i]
        su0, sv0 = int(p["size_u0"]), int(p["size_v0"])
        if su0 == 0 or sv0 == 0:
            continue
        v0, u0, v1, u1 = np.meshgrid(np.arange(sv0), np.arange(su0), np.arange(R), np.arange(R), indexing="ij")
        u, v = (u0 * R + u1).ravel(), (v0 * R + v1).ravel()
        x, y = patch2canvas(int(p["orientation"]), u, v, su0 * R, sv0 * R, int(p["u0"]) * R, int(p["v0"]) * R)
        keep = (b2p[y // R, x // R] == i + 1) & (occ[y, x] != 0)
        u, v, x, y = u[keep], v[keep], x[keep], y[keep]
        n, d1, mode = len(u), int(p["d1"]), int(p["projection_mode"])
        if use_plr:
            ownI, other, two, lo, nfill = _plr_pixels(p, plr, patch_base + i, geo, u, v, x, y, R)
            lay_of = (0, 100, 101)  # :826-828: 0, g_intermediateLayerIndex, g_intermediateLayerIndex + 1
        else:
            ownI, other, two, lo, nfill = _interleaved_pixels(p, i, geo, occ, b2p, x, y, R, st)
            lay_of = None
        emit1 = two & ~((other == ownI) & bool(P.remove_duplicate_points))  # :794-795
        per = 1 + emit1 + nfill
        tot = int(per.sum())
        pixi = np.repeat(np.arange(n), per)
        k = np.arange(tot) - np.repeat(np.cumsum(per) - per, per)  # position inside the pixel's run
        is0 = k == 0
        is1 = (k == 1) & emit1[pixi]
        fillk = k - 1 - emit1[pixi] + 1  # 1-based step of a fill point
        nn = np.where(is0, ownI[pixi], np.where(is1, other[pixi], lo[pixi] + fillk))
        Pn = np.zeros((tot, 3), np.int64)
        Pn[:, p["normal_axis"]] = nn
        Pn[:, p["tangent_axis"]] = u[pixi] * int(p["lod_x"]) + int(p["u1"])
        Pn[:, p["bitangent_axis"]] = v[pixi] * int(p["lod_y"]) + int(p["v1"])
        par = (x[pixi] + y[pixi]) & 1
        if lay_of is None:
            lay = np.where(is0, par, np.where(is1, par ^ 1, 100))  # :821-825, g_intermediateLayerIndex
        else:
            lay = np.where(is0, lay_of[0], np.where(is1, lay_of[1], lay_of[2]))
        pos.append(Pn.astype(np.int16))
        typ.append(btype[y[pixi], x[pixi]])
        part.append(np.full(tot, i, np.uint32))
        p2p.append(np.stack([x[pixi], y[pixi], lay], axis=1).astype(np.uint32))
        src.append(is0)
        if P.attribute_count > 0:
            c16 = np.stack([attribute[0, c, y[pixi], x[pixi]] for c in range(3)], axis=1).astype(np.uint16)
            col.append(np.where(is0[:, None], c16, 0).astype(np.uint16))

    def cat(parts, shape, dt):
        return np.concatenate(parts) if parts else np.zeros(shape, dt)
    n = sum(len(a) for a in pos)
    out = dict(positions=cat(pos, (0, 3), np.int16), boundary_types=cat(typ, (0,), np.uint16).astype(np.uint16),
               partition=cat(part, (0,), np.uint32), point_to_pixel=cat(p2p, (0, 3), np.uint32),
               colors16=cat(col, (0, 3), np.uint16) if P.attribute_count > 0 else np.zeros((n, 3), np.uint16),
               colors=np.zeros((n, 3), np.uint8))
    exact = np.ones(n, bool)
    is_src = cat(src, (0,), bool)
    if P.attribute_count > 0 and is_src.any() and (~is_src).any() and is_src.sum() >= 6:
        # transferColorWeight: 5-NN among the coded points, w = 1 / (d^2)^2, double sums in result order, truncation
        S, T = np.nonzero(is_src)[0], np.nonzero(~is_src)[0]
        sp, tp = out["positions"][S].astype(np.int64), out["positions"][T].astype(np.int64)
        _, nb = cKDTree(sp).query(tp, k=6)
        d2 = ((sp[nb] - tp[:, None, :]) ** 2).sum(axis=2)
        o = np.argsort(d2, axis=1, kind="stable")
        nb, d2 = np.take_along_axis(nb, o, 1), np.take_along_axis(d2, o, 1)
        strict = (np.diff(d2, axis=1) > 0).all(axis=1)  # no tie anywhere among the six nearest
        c = out["colors16"][S][nb[:, :5]].astype(np.float64)
        same = d2[:, 0] == 0  # result.dist( 0 ) <= 0.0001: the colour of that coded point (:2262, :2274-2276)
        w = 1.0 / (np.where(same[:, None], 1, d2[:, :5]).astype(np.float64) ** 2)
        acc, sw = np.zeros((len(T), 3)), np.zeros(len(T))
        for j in range(5):
            acc = acc + c[:, j] * w[:, j][:, None]
            sw = sw + w[:, j]
        res = (acc / sw[:, None]).astype(np.int64).astype(np.uint16)
        out["colors16"][T] = np.where(same[:, None], out["colors16"][S][nb[:, 0]], res)
        exact[T] = np.where(same, d2[:, 1] > 0, strict)
    out["colors16_exact"] = exact
    return out, b2p, occ


# ---------------------------------------------------------------------------------------------------------------
# PCCCodec::smoothPointCloudPostprocess (:52-147) + addGridCentroid (:980-998) + smoothPointCloudGrid / gridFiltering
# (:1000-1104)
# ---------------------------------------------------------------------------------------------------------------
def _cell_tables(cell_ids, values, partition):
    """per occupied cell: count, float32 sums accumulated in emission order, doSmooth (>= 2 distinct partitions)"""
    uniq, inv = np.unique(cell_ids, return_inverse=True)
    cnt = np.bincount(inv, minlength=len(uniq))
    sums = np.stack([np.bincount(inv, weights=values[:, k].astype(np.float64), minlength=len(uniq)) for k in range(3)], axis=1)
    assert sums.max(initial=0) < 2 ** 24, "float accumulator would leave the exact range (SURVEY App. A.3)"
    pmin = np.full(len(uniq), np.iinfo(np.int64).max)
    pmax = np.full(len(uniq), -1)
    np.minimum.at(pmin, inv, partition.astype(np.int64))
    np.maximum.at(pmax, inv, partition.astype(np.int64))
    return uniq, cnt, sums.astype(np.float32), pmin != pmax


def smooth_geometry(params, cloud):
    g = params.grid_size
    pos = cloud["positions"].astype(np.int64)
    typ = cloud["boundary_types"].copy()
    n = len(pos)
    if n == 0 or not params.flag_geometry_smoothing:
        return cloud
    if not params.grid_smoothing:  # :140-142
        return smooth_geometry_radius(params, cloud) if params.neighbor_count_smoothing > 0 and not params.pbf_enable else cloud
    w = (int(pos.max()) + g - 1) // g  # :68-79
    disth, th = max(g // 2, 1), g * w
    inside = ~((pos < disth).any(axis=1) | (th <= pos + disth).any(axis=1))  # :92-95
    W = w + 2
    cid = (pos[:, 0] // g) + (pos[:, 1] // g) * W + (pos[:, 2] // g) * W * W
    uniq, cnt, sums, multi = _cell_tables(cid[inside], pos[inside], cloud["partition"][inside])
    centre = (sums / cnt[:, None].astype(np.float32)).astype(np.float64)  # :135-137, one float division per component
    look = {int(c): k for k, c in enumerate(uniq)}
    out = cloud["positions"].copy()
    g2, hg = 2 * g, g // 2
    thr = int(params.threshold_smoothing)
    for i in np.nonzero(inside & (typ == 1))[0]:  # :1087
        Pp = pos[i]
        S = Pp // g - ((Pp % g) < hg)  # :1014-1017
        cells = [look.get(int((S[0] + dx) + (S[1] + dy) * W + (S[2] + dz) * W * W), -1)
                 for dz in (0, 1) for dy in (0, 1) for dx in (0, 1)]
        if not any(k >= 0 and multi[k] for k in cells):  # doSmooth && count (:1024)
            continue
        Wt = (Pp - S * g - hg) * 2 + 1  # :1034
        Q = g2 - Wt
        c4 = np.zeros(3)
        count = 0
        kk = 0
        for dz in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    wgt = int((Wt[0] if dx else Q[0]) * (Wt[1] if dy else Q[1]) * (Wt[2] if dz else Q[2]))
                    k = cells[kk]
                    kk += 1
                    v = centre[k] if k >= 0 else Pp.astype(np.float64)
                    c4 = c4 + v * float(wgt)  # :1050-1058
                    count += wgt * (int(cnt[k]) if k >= 0 else 0)
        c4 = c4 / float(g2 ** 3)
        count //= g2 ** 3  # int division; 0 is common (App. A.4): 0/0 = NaN, NaN >= x is false
        if count == 0:
            continue
        cen = c4 * float(count)
        e = Pp.astype(np.float64) * float(count) - cen
        dist2 = ((e[0] * e[0] + e[1] * e[1]) + e[2] * e[2]) / float(count) + 0.5  # :1093
        if dist2 >= float(max(thr, count) * 2):  # :1094
            out[i] = np.trunc(cen / float(count) + 0.5).astype(np.int64).astype(np.int16)  # :1095-1097
            typ[i] = 3
    res = dict(cloud)
    res["positions"], res["boundary_types"] = out, typ
    return res


# ---------------------------------------------------------------------------------------------------------------
# PCCCodec::smoothPointCloud (:1106-1157), the non-grid filter.  PCCKdTree::searchRadius (PCCKdTree.cpp:69-79) is nanoflann's
# radiusSearch (dist < radius2Smoothing) + std::sort with the vendored IndexDist_Sorter — by distance, equal distances by
# index (dependencies/nanoflann/nanoflann.hpp:193-200): a total order, so the neighbourhood is the first
# neighborCountSmoothing entries of the (distance, index)-sorted ball and no tree order is involved.
# ---------------------------------------------------------------------------------------------------------------
def smooth_geometry_radius(params, cloud):
    from scipy.spatial import cKDTree
    pos = cloud["positions"].astype(np.int64)
    part = cloud["partition"].astype(np.int64)
    typ = cloud["boundary_types"].copy()
    out = cloud["positions"].copy()
    r2, r2b, nmax = float(params.radius2_smoothing), float(params.radius2_boundary_detection), int(params.neighbor_count_smoothing)
    tree = cKDTree(pos.astype(np.float64))
    balls = tree.query_ball_point(pos.astype(np.float64), np.sqrt(r2) + 1e-9)  # superset; the exact test follows
    for i in range(len(pos)):
        idx = np.asarray(balls[i], np.int64)
        d = ((pos[idx] - pos[i]) ** 2).sum(axis=1)
        keep = d < r2  # RadiusResultSet::addPoint, :165-169
        idx, d = idx[keep], d[keep]
        order = np.lexsort((idx, d))[:nmax]  # sorted by distance then index, cut to num_results (PCCKdTree.cpp:75-76)
        idx, d = idx[order], d[order]
        cnt = len(idx)
        if cnt == 0 or not ((d <= r2b) & (part[idx] != part[i])).any():  # otherClusterPointCount, :1126-1128
            continue
        if typ[i] == 1:
            typ[i] = 2  # :1131-1133
        centroid = pos[idx].sum(axis=0).astype(np.float64)  # sums of int16 in double: exact
        e = centroid - float(cnt) * pos[i].astype(np.float64)
        dist = float(np.int64((e[0] * e[0] + e[1] * e[1] + e[2] * e[2]) + cnt / 2.0)) / float(cnt)  # :1136-1137
        if dist >= float(params.threshold_smoothing):  # :1141
            out[i] = np.trunc((centroid + float(cnt // 2)) / float(cnt)).astype(np.int64).astype(np.int16)  # :1138-1140
    res = dict(cloud)
    res["positions"], res["boundary_types"] = out, typ
    return res


# ---------------------------------------------------------------------------------------------------------------
# PCCCodec::colorSmoothing (:149-236) + addGridColorCentroid (:1159-1180) + gridFilteringColor (:1182-1266) +
# smoothPointCloudColorLC (:1268-1306); median / mean: PCCCodec.h:271-285
# ---------------------------------------------------------------------------------------------------------------
def smooth_color(params, cloud):
    g = params.occupancy_precision  # the colour grid uses occupancyPrecision (:152)
    pcmax = 1 << params.geometry_bitdepth_3d
    w = pcmax // g
    pos = cloud["positions"].astype(np.int64)
    typ = cloud["boundary_types"]
    col = cloud["colors16"].astype(np.int64)
    n = len(pos)
    if n == 0:
        return cloud
    disth, hg, g2 = max(g // 2, 1), g // 2, 2 * g
    ok = (pos >= 0).all(axis=1) & ((pos // g) < w).all(axis=1)  # :212 guard
    cid = (pos[:, 0] // g) + (pos[:, 1] // g) * w + (pos[:, 2] // g) * w * w
    uniq, cnt, sums, multi = _cell_tables(cid[ok], col[ok], cloud["partition"][ok])
    mean = sums.astype(np.float64) / cnt[:, None].astype(np.float64)  # :1225
    look = {int(c): k for k, c in enumerate(uniq)}
    # per-cell luma lists in emission order (colorSmoothingLum_, :1179)
    order = np.argsort(cid[ok], kind="stable")
    lum_sorted = col[ok][order, 0]
    starts = np.concatenate([[0], np.cumsum(cnt)])
    var_thr, diff_thr = params.threshold_color_variation * 256.0, params.threshold_color_difference * 256.0

    def gate(k):  # |trunc(mean - median)| > variation * 256, int abs(int) on a truncated double (App. A.9)
        if cnt[k] <= 1:
            return False
        lum = np.sort(lum_sorted[starts[k]:starts[k + 1]])
        m = len(lum)
        med = (lum[m // 2] + lum[m // 2 - 1]) / 2.0 if m % 2 == 0 else float(lum[m // 2])
        mn = float(lum.sum()) / m
        return abs(int(mn - med)) > var_thr
    gates = {}
    out = cloud["colors16"].copy()
    inside = ~((pos < disth).any(axis=1) | (pcmax <= pos + disth).any(axis=1))  # :1280-1283
    for i in np.nonzero(inside & (typ == 1))[0]:  # :1288
        Pp = pos[i]
        S = Pp // g - ((Pp % g) < hg)  # :1197-1199
        cells = [look.get(int((S[0] + dx) + (S[1] + dy) * w + (S[2] + dz) * w * w), -1)
                 if 0 <= S[0] + dx < w and 0 <= S[1] + dy < w and 0 <= S[2] + dz < w else -1
                 for dz in (0, 1) for dy in (0, 1) for dx in (0, 1)]
        if not any(k >= 0 and multi[k] for k in cells):  # :1204
            continue
        cur = col[i].astype(np.float64)
        Wt = (Pp - S * g - hg) * 2 + 1
        Q = g2 - Wt
        c3 = []
        keep_own = False
        Y0 = 0.0
        for kk, k in enumerate(cells):  # :1218-1251
            if k >= 0:
                if k not in gates:
                    gates[k] = gate(k)
                d = mean[k].copy()
                if kk == 0:
                    if gates[k]:
                        keep_own = True
                        break
                else:
                    own = abs(int(Y0 - d[0])) > diff_thr  # :1238
                    if gates[k]:
                        own = True
                    if own:
                        d = cur.copy()
            else:
                d = cur.copy()
            if kk == 0:
                Y0 = d[0]
            c3.append(d)
        if keep_own:
            continue
        c4 = np.zeros(3)
        kk = 0
        for dz in (0, 1):
            for dy in (0, 1):
                for dx in (0, 1):
                    wgt = float(int((Wt[0] if dx else Q[0]) * (Wt[1] if dy else Q[1]) * (Wt[2] if dz else Q[2])))
                    c4 = c4 + c3[kk] * wgt
                    kk += 1
        res = np.trunc(c4 / float(g2 ** 3) + 0.5)  # :1262, :1294-1295
        dist = abs(int(res[0] - cur[0])) * 10.0 / 256.0  # :1299
        if dist >= params.threshold_color_smoothing:
            out[i] = res.astype(np.int64).astype(np.uint16)
    r = dict(cloud)
    r["colors16"] = out
    return r


# ---------------------------------------------------------------------------------------------------------------
# PCCPointSet3::convertYUV16ToRGB8 / copyRGB16ToRGB8 (PccLibCommon/include/PCCPointSet.h:121-166)
# ---------------------------------------------------------------------------------------------------------------
def to_rgb8(params, cloud):
    c = cloud["colors16"].astype(np.float64)
    out = dict(cloud)
    if params.attribute_count == 0:
        out["colors"] = np.full((len(c), 3), 127, np.uint8)
        return out
    if params.attribute_rgb444:
        out["colors"] = cloud["colors16"].astype(np.uint8)
        return out
    wgt = 1.0 / 65535.0
    y1 = np.clip(wgt * c[:, 0], 0.0, 1.0)
    u1 = np.clip(wgt * (c[:, 1] - 32768.0), -0.5, 0.5)
    v1 = np.clip(wgt * (c[:, 2] - 32768.0), -0.5, 0.5)
    r = y1 + 1.57480 * v1
    g = (y1 - 0.18733 * u1) - 0.46813 * v1
    b = y1 + 1.85563 * u1

    def rnd(a):  # round() = half away from zero, then clip to [0, 255]
        a = a * 255.0
        return np.clip(np.where(a >= 0, np.floor(a + 0.5), np.ceil(a - 0.5)), 0, 255)
    out["colors"] = np.stack([rnd(r), rnd(g), rnd(b)], axis=1).astype(np.uint8)
    return out


# ---------------------------------------------------------------------------------------------------------------
# PCCPointSet3::removeDuplicate( out, dropDuplicates ) (PccLibCommon/source/PCCPointSet.cpp:169-218)
# ---------------------------------------------------------------------------------------------------------------
def remove_duplicates(positions, colors, drop=2):
    p = positions.astype(np.int64)
    key = ((p[:, 0] + 32768) << 32) | ((p[:, 1] + 32768) << 16) | (p[:, 2] + 32768)
    order = np.argsort(key, kind="stable")  # nested std::map<float,...>: lexicographic (x, y, z); lists in index order
    ks = key[order]
    first = np.concatenate([[True], ks[1:] != ks[:-1]])
    starts = np.nonzero(first)[0]
    out_p = positions[order][starts]
    if colors is None:
        return out_p, None
    c = colors[order].astype(np.int64)
    if drop == 1:
        return out_p, colors[order][starts]
    counts = np.diff(np.concatenate([starts, [len(ks)]]))
    sums = np.add.reduceat(c, starts, axis=0)
    return out_p, (sums // counts[:, None]).astype(np.uint8)  # size_t division (:197-199)


# ---------------------------------------------------------------------------------------------------------------
# QualityMetrics::compute — D1 (point-to-point) and colour only (PccLibMetrics/source/PCCMetrics.cpp:75-231)
# ---------------------------------------------------------------------------------------------------------------
def _yuv709(rgb):  # convertRGBtoYUVBT709, :50-55 (double arithmetic, each channel narrowed to float)
    r, g, b = (rgb[:, k].astype(np.float64) for k in range(3))
    return np.stack([((0.2126 * r + 0.7152 * g) + 0.0722 * b) / 255.0,
                     ((-0.1146 * r - 0.3854 * g) + 0.5000 * b) / 255.0 + 0.5000,
                     ((0.5000 * r - 0.4542 * g) - 0.0458 * b) / 255.0 + 0.5000], axis=1).astype(np.float32)


def quality_d1_colour(posA, colA, posB, colB):
    """one direction A -> B; the tie set is every point of B at the nearest distance, colours averaged (neighborsProc 1)"""
    from scipy.spatial import cKDTree
    tree = cKDTree(posB.astype(np.float64))
    d, _ = tree.query(posA.astype(np.float64), k=1)
    d2 = np.rint(d * d).astype(np.int64)
    sse_c2c = float(d2.sum())
    yA = _yuv709(colA)
    sse_col = np.zeros(3)
    ties = tree.query_ball_point(posA.astype(np.float64), np.sqrt(d2) + 1e-6)
    for i, t in enumerate(ties):
        pb = posB[t].astype(np.int64) - posA[i].astype(np.int64)
        t = np.array(t)[(pb * pb).sum(axis=1) == d2[i]]
        assert 0 < len(t) <= 30
        s = colB[t].astype(np.int64).sum(axis=0)
        rgb = np.floor(s / len(t) + 0.5).astype(np.uint8)[None, :]  # round() of a non-negative double
        yB = _yuv709(rgb)[0]
        diff = (yA[i] - yB).astype(np.float32)
        sse_col += (diff * diff).astype(np.float32).astype(np.float64)  # powf( float, 2 ) then summed in double
    return sse_c2c, sse_col, len(posA)


class Port:
    """same calling convention as oracle.checker backends for the stages it restates"""

    def run_gof(self, gof, stages=("reconstruct", "smooth_geometry", "smooth_color", "rgb8")):
        P = gof.params
        out = []
        for f in range(gof.n_frames):
            patches = gof.patches[gof.patch_offset[f]:gof.patch_offset[f + 1]]
            raws = gof.raw_patches[gof.raw_offset[f]:gof.raw_offset[f + 1]] if gof.raw_patches is not None else None
            aux_g = gof.aux_geometry[f] if getattr(gof, "aux_geometry", None) is not None else None
            aux_a = gof.aux_attribute[f] if getattr(gof, "aux_attribute", None) is not None else None
            cloud, b2p, occ = reconstruct_frame(P, gof.occupancy[f], gof.geometry[f], gof.attribute[f], patches,
                                                getattr(gof, "plr", None), int(gof.patch_offset[f]), raws, aux_g, aux_a)
            exact = cloud.pop("colors16_exact", None)  # pixel interleaving: colours that do not hinge on kd-tree tie order
            snaps = {"reconstruct": cloud, "block_to_patch": b2p, "occupancy": occ, "colors16_exact": exact}
            if P.apply_geo_smoothing and P.flag_geometry_smoothing:
                if P.grid_smoothing or P.neighbor_count_smoothing > 0:  # (> 0 without the grid: the encoder-side call)
                    cloud = smooth_geometry(P, cloud)
                snaps["smooth_geometry"] = cloud
                if P.attribute_count > 0 and P.attr_transfer_filter_type != 0:
                    raise NotImplementedError("transferColors16bitBP is not restated in oracle_np (use oracle/_ref)")
            if P.attribute_count > 0:
                if P.apply_attr_smoothing and P.flag_color_smoothing:
                    cloud = smooth_color(P, cloud)
                snaps["smooth_color"] = cloud
            cloud = to_rgb8(P, cloud)
            snaps["rgb8"] = cloud
            out.append(snaps)
        return out


# ---------------------------------------------------------------------------------------------------------------
# PCCInternalColorConverter<T>::convertYUV420ToYUV444 (PccLibColorConverter/source/PCCInternalColorConverter.cpp:
# 456-486) as PCCVideoDecoder runs it after the attribute video is decoded (PccLibDecoder/source/PCCVideoDecoder.cpp:
# 125-146, :365): YUVtoFloatYUV (:596-611), upsampling (:669-695) with upsamplingVertical0/1 and
# upsamplingHorizontal0/1 (include/PCCInternalColorConverter.h:187-249), floatYUVToYUV (:582-594).
# All filter arithmetic is sequential single-precision (np.float32), the sample scaling goes through double.
# ---------------------------------------------------------------------------------------------------------------
# g_filter420to444 (PCCInternalColorConverter.cpp:297-337): (horizontal0, vertical0, horizontal1, vertical1), shift 8
UPSAMPLING_FILTERS = [
    ((0, 256), (-8, 64, 216, -16), (-16, 144, 144, -16), (-16, 216, 64, -8)),                                    # UF_F0
    ((0, 256), (0, -16, 56, 240, -32, 8), (-16, 144, 144, -16), (8, -32, 240, 56, -16, 0)),                      # UF_FV
    ((0, 256), (-6, 58, 222, -18), (-16, 144, 144, -16), (-18, 222, 58, -6)),                                    # UF_GS
    ((0, 256), (2, -18, 70, 228, -34, 8), (6, -34, 156, 156, -34, 6), (8, -34, 228, 70, -18, 2)),                # UF_LS3
    ((0, 256), (-1, 8, -23, 72, 229, -39, 14, -4), (-3, 15, -43, 159, 159, -43, 15, -3),
     (-4, 14, -39, 229, 72, -23, 8, -1)),                                                                        # UF_LS4
    ((0, 256), (3, -16, 67, 227, -32, 7), (21, -52, 159, 159, -52, 21), (7, -32, 227, 67, -16, 3)),              # UF_TM
    ((0, 256), (1, -5, 12, -27, 74, 230, -41, 18, -8, 2), (2, -8, 21, -47, 160, 160, -47, 21, -8, 2),
     (2, -8, 18, -41, 230, 74, -27, 12, -5, 1)),                                                                 # UF_LS5
    ((0, 256), (0, 3, -7, 14, -29, 75, 230, -43, 20, -10, 5, -2), (-1, 5, -12, 24, -49, 161, 161, -49, 24, -12, 5, -1),
     (-2, 5, -10, 20, -43, 230, 75, -29, 14, -7, 3, 0)),                                                         # UF_LS6
]


def image_set(src, shift):
    """PCCImage<T,3>::set (PccLibCommon/include/PCCImage.h:97-138): decoder samples -> stored uint16 samples"""
    if shift <= 0:
        return src.astype(np.uint16)
    v = ((src.astype(np.int32) + (1 << (shift - 1))) >> shift) & 0xFFFF
    return np.clip(v, 0, (1 << (10 - shift)) - 1).astype(np.uint16)


def _yuv_to_float(src, chroma, nbyte):  # YUVtoFloatYUV :596-611
    offset = (128 if nbyte == 1 else 512) if chroma else 0
    weight = 1.0 / (255.0 if nbyte == 1 else 1023.0)
    f = (weight * (src.astype(np.int32) - offset).astype(np.float64)).astype(np.float32)
    return np.clip(f, np.float32(-0.5 if chroma else 0.0), np.float32(0.5 if chroma else 1.0))


def _float_to_yuv16(f, chroma):  # floatYUVToYUV :582-594 with nbyte = 2
    x = (65535.0 * f.astype(np.float64) + (32768.0 if chroma else 0.0)).astype(np.float32)
    t = np.trunc(x)  # std::round(float): half away from zero, without forming x + 0.5f
    r = (t + np.where(np.abs(x - t) >= np.float32(0.5), np.sign(x), np.float32(0))).astype(np.float32)
    return np.clip(r, np.float32(0), np.float32(65535)).astype(np.uint16)


def _fir(im, taps, axis, shift_index):
    """value = sum_t taps[t] * im[clamp(i0 + t - position)] accumulated in float32 in tap order, * 1/256;
    i0 = index + shift_index (0 for the phase-0 filters, 1 for the phase-1 filters)"""
    n = im.shape[axis]
    position = (len(taps) + 1) >> 1
    idx = np.arange(n)
    value = np.zeros(im.shape, np.float32)
    for t, c in enumerate(taps):
        src = np.clip(idx + shift_index + t - position, 0, n - 1)
        value = (value + np.float32(c) * np.take(im, src, axis=axis)).astype(np.float32)
    return ((value + np.float32(0.0)) * np.float32(1.0 / 256.0)).astype(np.float32)


def _upsample(ch, filt):  # upsampling :669-695
    h0, v0, h1, v1 = UPSAMPLING_FILTERS[filt]
    hh, ww = ch.shape
    temp = np.empty((2 * hh, ww), np.float32)
    temp[0::2] = _fir(ch, v0, 0, 0)
    temp[1::2] = _fir(ch, v1, 0, 1)
    out = np.empty((2 * hh, 2 * ww), np.float32)
    out[:, 0::2] = _fir(temp, h0, 1, 0)
    out[:, 1::2] = _fir(temp, h1, 1, 1)
    return out


def yuv420_to_yuv444(y, u, v, bitdepth, filt):
    """y [H][W], u / v [H/2][W/2] decoded samples (8 or 10 bit) -> [3][H][W] uint16 (16-bit 4:4:4)"""
    nbyte = 1 if bitdepth == 8 else 2
    out = np.empty((3,) + y.shape, np.uint16)
    out[0] = _float_to_yuv16(_yuv_to_float(y, False, nbyte), False)
    out[1] = _float_to_yuv16(_upsample(_yuv_to_float(u, True, nbyte), filt), True)
    out[2] = _float_to_yuv16(_upsample(_yuv_to_float(v, True, nbyte), filt), True)
    return out


def convert_bitdepth(plane, bits_in, bits_out, msb_align):
    """PCCImage<T, N>::convertBitdepth (PccLibCommon/source/PCCImage.cpp:258-299): the result is stored back into the
    sample type T of `plane` (uint8 / uint16), i.e. truncated to it."""
    a = np.asarray(plane)
    diff = int(bits_in) - int(bits_out)
    v = a.astype(np.int64)
    if diff >= 0:
        v = (v >> diff) if msb_align else np.minimum(v, (1 << bits_out) - 1)
    elif msb_align:
        v = v << (-diff)
    return v.astype(a.dtype)  # C conversion to T: modulo 2^bits
