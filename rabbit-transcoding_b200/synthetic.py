"""Deterministic synthetic "decoded V-PCC GOF" generator.

Produces exactly what PCCDecoder holds after the three video sub-bitstreams are decoded and before the
per-frame reconstruction loop starts (PccLibDecoder/source/PCCDecoder.cpp:330): occupancy / geometry D0,D1 /
4:4:4 16-bit attribute frames, the per-frame patch tables, and the GeneratePointCloudParameters.  All
integer, seeded; the CUDA path, the CPU restatement and the unmodified reference consume the same arrays.

Shape: a closed surface made of a union of ellipsoids ("humanoid": torso, head, limbs) inside a 2^bitdepth
cube, projected orthographically on the 6 axis-aligned planes (view ids 0..5 of PCCPatch::setViewId,
PCCPatch.cpp:111-137), segmented by dominant normal, cut into patches on 16x16 blocks and packed into a
W-wide atlas with an occupancy-aware first-fit, as the encoder does.  Coding noise (+-1 on a fraction of the
depth samples) creates smoothing candidates.  The noise-free voxels with RGB + analytic normals form the
"source" cloud for the metrics.
"""
import numpy as np

from . import abi

# PCCPatch::setViewId: viewId -> (normal, tangent, bitangent, projectionMode)
VIEW_AXES = {0: (0, 2, 1, 0), 1: (1, 2, 0, 0), 2: (2, 0, 1, 0), 3: (0, 2, 1, 1), 4: (1, 2, 0, 1), 5: (2, 0, 1, 1)}
# enum PCCPatchOrientation (PCCBitstreamCommon.h:120-130): 0 DEFAULT 1 SWAP 2 ROT90 3 ROT180 4 ROT270 5 MIRROR
# 6 MROT90 7 MROT180 8 MROT270; these exchange U and V on the canvas (PCCPatch.h:63-70):
_SWITCHED = (1, 2, 4, 6, 8)


def default_params(width, height, bitdepth=10, occupancy_precision=4):
    """CTC C2 lossy defaults: cfg/common/ctc-common.cfg:22,25,57-60, Rec-1 (PCCDecoderParameters.cpp:125-134)."""
    p = abi.Params()
    p.width, p.height = width, height
    p.occupancy_resolution = 16
    p.occupancy_precision = occupancy_precision
    p.threshold_lossy_om = 0
    p.map_count_minus1 = 1
    p.absolute_d1 = 1
    p.remove_duplicate_points = 1
    p.geometry_bitdepth_3d = bitdepth
    p.attribute_count = 1
    p.attribute_rgb444 = 0
    p.flag_geometry_smoothing = 1
    p.grid_smoothing = 1
    p.grid_size = 8
    p.threshold_smoothing = 64.0
    p.apply_geo_smoothing = 1
    p.attr_transfer_filter_type = 1
    p.flag_color_smoothing = 1
    p.apply_attr_smoothing = 1
    p.threshold_color_smoothing = 10.0
    p.threshold_color_difference = 10.0
    p.threshold_color_variation = 6.0
    p.surface_thickness = 4
    return p


class SyntheticGOF:
    """Container of one generated GOF (host numpy arrays + the C structs pointing at them)."""

    def __init__(self):
        self.params = None
        self.n_frames = 0
        self.occupancy = None   # u8  [F][H/p][W/p]
        self.geometry = None    # u16 [F][M][H][W]
        self.attribute = None   # u16 [F][M][3][H][W]
        self.patches = None     # PATCH_DTYPE [total]
        self.patch_offset = None
        self.eom_patches = None
        self.eom_offset = None
        self.eom_members = None
        self.raw_patches = None
        self.raw_offset = None
        self.aux_geometry = None   # u16 [F][aux_height][aux_width]     (params.use_aux_separate_video, make_aux_video)
        self.aux_attribute = None  # u16 [F][3][aux_height][aux_width]
        self.plr = None         # point local reconstruction tables (make_plr)
        self.sources = []       # per frame: dict(positions i16[n,3], colors u8[n,3], normals f32[n,3])

    def frames_struct(self):
        f = abi.Frames()
        f.occupancy = abi.ptr(self.occupancy)
        f.geometry = abi.ptr(self.geometry)
        f.attribute = abi.ptr(self.attribute) if self.attribute is not None else None
        f.aux_geometry = abi.ptr(self.aux_geometry) if self.aux_geometry is not None else None
        f.aux_attribute = abi.ptr(self.aux_attribute) if self.aux_attribute is not None else None
        return f

    def atlas_struct(self):
        a = abi.Atlas()
        a.patches = abi.ptr(self.patches)
        a.patch_offset = abi.ptr(self.patch_offset)
        a.eom_patches = abi.ptr(self.eom_patches) if self.eom_patches is not None else None
        a.eom_offset = abi.ptr(self.eom_offset) if self.eom_offset is not None else None
        a.eom_members = abi.ptr(self.eom_members) if self.eom_members is not None else None
        a.raw_patches = abi.ptr(self.raw_patches) if self.raw_patches is not None else None
        a.raw_offset = abi.ptr(self.raw_offset) if self.raw_offset is not None else None
        return a

    def input_bytes(self):
        n = self.occupancy.nbytes + self.geometry.nbytes + self.patches.nbytes
        if self.attribute is not None:
            n += self.attribute.nbytes
        return n


def _humanoid(cube, scale, rng, t):
    """list of (centre[3], radii[3]) in voxel units; t = frame phase for motion"""
    s = cube * scale
    cx, cz = cube * 0.5, cube * 0.5
    sway = 0.01 * cube * np.sin(0.37 * t)
    step = 0.015 * cube * np.sin(0.61 * t + 0.5)
    parts = [
        ((cx + sway, 0.63 * cube, cz), (0.19 * s, 0.26 * s, 0.13 * s)),                     # torso
        ((cx + sway, 0.63 * cube + 0.36 * s, cz + step * 0.2), (0.09 * s, 0.11 * s, 0.10 * s)),   # head
        ((cx - 0.26 * s + sway, 0.60 * cube, cz + step), (0.06 * s, 0.24 * s, 0.06 * s)),    # arm L
        ((cx + 0.26 * s + sway, 0.60 * cube, cz - step), (0.06 * s, 0.24 * s, 0.06 * s)),    # arm R
        ((cx - 0.10 * s, 0.63 * cube - 0.50 * s, cz - step), (0.08 * s, 0.30 * s, 0.08 * s)),  # leg L
        ((cx + 0.10 * s, 0.63 * cube - 0.50 * s, cz + step), (0.08 * s, 0.30 * s, 0.08 * s)),  # leg R
        ((cx, 0.63 * cube - 0.18 * s, cz), (0.17 * s, 0.12 * s, 0.12 * s)),                 # hips
    ]
    jit = rng.uniform(-0.004, 0.004, size=(len(parts), 3)) * cube
    return [(np.array(c) + j, np.array(r)) for (c, r), j in zip(parts, jit)]


def _view_depth(parts, cube, normal, tangent, bitangent, mode):
    """orthographic depth map along `normal`; image axes (v=bitangent rows, u=tangent cols).
    returns depth (float, nan where empty), part index map."""
    depth = np.full((cube, cube), np.nan, dtype=np.float64)
    owner = np.full((cube, cube), -1, dtype=np.int32)
    for k, (c, r) in enumerate(parts):
        u_lo, u_hi = int(max(0, np.floor(c[tangent] - r[tangent]))), int(min(cube, np.ceil(c[tangent] + r[tangent]) + 1))
        v_lo, v_hi = int(max(0, np.floor(c[bitangent] - r[bitangent]))), int(min(cube, np.ceil(c[bitangent] + r[bitangent]) + 1))
        if u_hi <= u_lo or v_hi <= v_lo:
            continue
        uu = (np.arange(u_lo, u_hi) - c[tangent]) / r[tangent]
        vv = (np.arange(v_lo, v_hi) - c[bitangent]) / r[bitangent]
        q = 1.0 - uu[None, :] ** 2 - vv[:, None] ** 2
        inside = q > 0
        h = np.sqrt(np.where(inside, q, 0.0)) * r[normal]
        d = c[normal] - h if mode == 0 else c[normal] + h
        sub = depth[v_lo:v_hi, u_lo:u_hi]
        osub = owner[v_lo:v_hi, u_lo:u_hi]
        if mode == 0:
            better = inside & (np.isnan(sub) | (d < sub))
        else:
            better = inside & (np.isnan(sub) | (d > sub))
        sub[better] = d[better]
        osub[better] = k
    return depth, owner


def _split_rects(mask, max_blocks, rng, R=16):
    """cut the occupied area of `mask` into block-aligned rectangles (patch bounding boxes)."""
    out = []
    ys, xs = np.nonzero(mask)
    if len(ys) == 0:
        return out
    stack = [(xs.min(), ys.min(), xs.max() + 1, ys.max() + 1)]
    while stack:
        x0, y0, x1, y1 = stack.pop()
        sub = mask[y0:y1, x0:x1]
        if not sub.any():
            continue
        yy, xx = np.nonzero(sub)
        x0, x1, y0, y1 = x0 + xx.min(), x0 + xx.max() + 1, y0 + yy.min(), y0 + yy.max() + 1
        w, h = x1 - x0, y1 - y0
        lim = int(rng.integers(max(2, max_blocks // 2), max_blocks + 1)) * R
        if w <= lim and h <= lim:
            out.append((x0, y0, x1, y1))
            continue
        if w >= h:
            cut = x0 + int(rng.integers(max(1, w // (3 * R)), max(2, 2 * w // (3 * R) + 1))) * R
            cut = min(max(cut, x0 + R), x1 - 1)
            stack.append((x0, y0, cut, y1))
            stack.append((cut, y0, x1, y1))
        else:
            cut = y0 + int(rng.integers(max(1, h // (3 * R)), max(2, 2 * h // (3 * R) + 1))) * R
            cut = min(max(cut, y0 + R), y1 - 1)
            stack.append((x0, y0, x1, cut))
            stack.append((x0, cut, x1, y1))
    return out


def _correlate_valid(grid, cm):
    """number of colliding blocks for every placement of `cm` inside `grid` ('valid' correlation)"""
    from scipy.signal import correlate2d
    return correlate2d(grid.astype(np.int32), cm.astype(np.int32), mode="valid")


def _patch2canvas_arrays(orient, U, V, su, sv):
    """vectorised PCCPatch::patch2Canvas (PCCPatch.cpp:192-251) relative to the patch origin, pixel units.
    U,V patch-local pixel coords; su, sv = patch size in pixels (sizeU0*R, sizeV0*R)."""
    if orient == 0:
        return U, V
    if orient in (1, 8):
        return V, U
    if orient == 2:
        return sv - 1 - V, U
    if orient == 3:
        return su - 1 - U, sv - 1 - V
    if orient == 4:
        return V, su - 1 - U
    if orient == 5:
        return su - 1 - U, V
    if orient == 6:
        return sv - 1 - V, su - 1 - U
    if orient == 7:
        return U, sv - 1 - V
    raise ValueError(orient)


def _yuv16_field(P, rng_noise, amp):
    """smooth 3-D colour field -> 16-bit YUV (8-bit content << 8), P float [...,3]"""
    x, y, z = P[..., 0], P[..., 1], P[..., 2]
    Y = 120 + 70 * np.sin(x * 0.021 + 0.3) * np.cos(y * 0.013) + 30 * np.sin(z * 0.05)
    Cb = 128 + 50 * np.sin(y * 0.017 + z * 0.011)
    Cr = 128 + 50 * np.cos(x * 0.015 - y * 0.009)
    yuv = np.stack([Y, Cb, Cr], axis=-1)
    if rng_noise is not None and amp > 0:
        yuv = yuv + rng_noise.normal(0, amp, size=yuv.shape)
    return np.clip(np.rint(yuv * 256.0), 0, 65535).astype(np.uint16)


def yuv16_to_rgb8(c16):
    """PCCPointSet3::convertYUV16ToRGB8 (PCCPointSet.h:133-166) in numpy float64 (source-cloud colours)."""
    c = c16.astype(np.float64)
    w = 1.0 / 65535.0
    y1 = np.clip(w * c[..., 0], 0, 1)
    u1 = np.clip(w * (c[..., 1] - 32768.0), -0.5, 0.5)
    v1 = np.clip(w * (c[..., 2] - 32768.0), -0.5, 0.5)
    r = y1 + 1.57480 * v1
    g = y1 - 0.18733 * u1 - 0.46813 * v1
    b = y1 + 1.85563 * u1

    def rnd(a):  # C round(): half away from zero
        return np.clip(np.where(a >= 0, np.floor(a * 255 + 0.5), np.ceil(a * 255 - 0.5)), 0, 255)
    return np.stack([rnd(r), rnd(g), rnd(b)], axis=-1).astype(np.uint8)


def generate_gof(n_frames=1, bitdepth=10, width=1280, occupancy_precision=4, scale=1.0, seed=0x0AB817,
                 noise_fraction=0.10, color_noise=6.0, max_patch_blocks=12, orientations=(0, 1),
                 eom=False, raw_points=0, map_count=2, precedence_reverse=False, min_height_blocks=0,
                 with_sources=True, color_smoothing=True, geometry_smoothing=True, transfer_filter=1,
                 height_blocks=None, absolute_d1=True, frame_offset=0, max_depth=None):
    """Generate one GOF.  `scale` sizes the body relative to the cube (1.0 ~ vox10-like point counts at
    bitdepth 10: ~0.4 M occupied pixels, ~0.8 M points).  `height_blocks` forces the atlas height (all
    frames of a GOF share W x H as the video does); otherwise H = tallest packing over the GOF."""
    R = 16
    cube = 1 << bitdepth
    p_occ = 1 if eom else occupancy_precision
    Wb = width // R
    M = map_count
    frames = []
    fo = frame_offset  # frame f of this call is frame fo + f of the sequence (seeds and motion phase)
    for f in range(n_frames):
        rng = np.random.default_rng([seed, fo + f])
        parts = _humanoid(cube, scale, rng, float(fo + f))
        # ---- per-view depth maps and normal-based segmentation ----
        views = []
        for vid, (nrm, tan, bit, mode) in VIEW_AXES.items():
            depth, owner = _view_depth(parts, cube, nrm, tan, bit, mode)
            valid = (owner >= 0) & (np.rint(np.nan_to_num(depth, nan=-1.0)) >= 0) & (np.rint(np.nan_to_num(depth, nan=-1.0)) <= cube - 1)
            dz = np.rint(np.where(valid, depth, 0)).astype(np.int64)
            # surface normal of the owning ellipsoid at the hit point
            vv, uu = np.mgrid[0:cube, 0:cube]
            cen = np.array([pt[0] for pt in parts])
            rad = np.array([pt[1] for pt in parts])
            ow = np.where(valid, owner, 0)
            pos = np.zeros((cube, cube, 3))
            pos[..., nrm] = np.where(valid, depth, 0)
            pos[..., tan] = uu
            pos[..., bit] = vv
            g = (pos - cen[ow]) / (rad[ow] ** 2)
            gn = np.linalg.norm(g, axis=-1, keepdims=True)
            g = g / np.maximum(gn, 1e-12)
            dom = (np.abs(g[..., nrm]) >= np.abs(g[..., tan]) - 1e-9) & (np.abs(g[..., nrm]) >= np.abs(g[..., bit]) - 1e-9)
            sel = valid & dom
            views.append(dict(vid=vid, nrm=nrm, tan=tan, bit=bit, mode=mode, depth=dz, valid=valid, sel=sel,
                              normal=g.astype(np.float32)))
        # ---- cut into patches ----
        plist = []
        for vw in views:
            for (x0, y0, x1, y1) in _split_rects(vw["sel"], max_patch_blocks, rng, R):
                x0a, y0a = (x0 // R) * R if rng.random() < 0.5 else x0, (y0 // R) * R if rng.random() < 0.5 else y0
                su0, sv0 = -(-(x1 - x0a) // R), -(-(y1 - y0a) // R)
                # keep the patch inside the cube image
                if x0a + su0 * R > cube:
                    x0a = cube - su0 * R
                if y0a + sv0 * R > cube:
                    y0a = cube - sv0 * R
                m = np.zeros((sv0 * R, su0 * R), bool)
                m[:, :] = vw["sel"][y0a:y0a + sv0 * R, x0a:x0a + su0 * R]
                # a pixel belongs to this patch only inside the cut rectangle
                yy, xx = np.mgrid[0:sv0 * R, 0:su0 * R]
                m &= (xx + x0a >= x0) & (xx + x0a < x1) & (yy + y0a >= y0) & (yy + y0a < y1)
                if not m.any():
                    continue
                plist.append(dict(view=vw, u1=x0a, v1=y0a, su0=su0, sv0=sv0, mask=m))
        order = rng.permutation(len(plist))
        # biggest first, like the encoder's packing order
        order = sorted(order, key=lambda i: -(plist[i]["su0"] * plist[i]["sv0"]))
        plist = [plist[i] for i in order]
        for pt in plist:
            pt["orient"] = int(orientations[int(rng.integers(0, len(orientations)))])
        frames.append(dict(parts=parts, views=views, patches=plist, rng=rng))

    # ---- packing: occupancy-aware first fit on the block grid ----
    Hb_needed = 0
    for fr in frames:
        rows_cap = 4 * (cube // R) + 64
        grid = np.zeros((rows_cap, Wb), bool)
        for pt in fr["patches"]:
            su0, sv0, o = pt["su0"], pt["sv0"], pt["orient"]
            bm = pt["mask"].reshape(sv0, R, su0, R).any(axis=(1, 3))  # [sv0][su0] occupied blocks (patch space)
            VB, UB = np.mgrid[0:sv0, 0:su0]
            xb, yb = _patch2canvas_arrays(o, UB, VB, su0, sv0)
            cw, ch = (sv0, su0) if o in _SWITCHED else (su0, sv0)
            cm = np.zeros((ch, cw), bool)
            cm[yb, xb] = bm
            if cw > Wb:
                raise ValueError("patch wider than the atlas")
            placed = False
            # first position in raster order where no occupied block collides (encoder-style packing)
            top = 0
            while not placed and top < rows_cap - ch:
                bot = min(rows_cap, top + ch + 64)
                # the decoder gives a block to the LAST patch whose bounding box covers it (PCCCodec.cpp:1725-1763), so
                # a patch may sit in the holes of earlier bounding boxes but its own box must not cover occupied blocks
                hit = _correlate_valid(grid[top:bot], np.ones_like(cm))
                ys, xs = np.nonzero(hit == 0)
                if len(ys):
                    y, x = int(ys[0]) + top, int(xs[0])
                    grid[y:y + ch, x:x + cw] |= cm
                    pt["u0"], pt["v0"] = x, y
                    placed = True
                else:
                    top += 64
            if not placed:
                raise RuntimeError("packing failed")
        # the whole bounding box of every patch must lie inside the canvas (PCCPatch.cpp:237-245 exits otherwise)
        fr["Hb"] = max([1] + [pt["v0"] + (pt["su0"] if pt["orient"] in _SWITCHED else pt["sv0"]) for pt in fr["patches"]])
        Hb_needed = max(Hb_needed, fr["Hb"])

    # extra room for EOM / raw rectangles
    eom_rows = raw_rows = 0
    if eom:
        eom_rows = 0  # fixed after counting below
    if raw_points:
        raw_rows = -(-(3 * raw_points) // (Wb * R * R))
    Hb = max(Hb_needed, min_height_blocks)

    # ---- rasterise frames ----
    gof = SyntheticGOF()
    per_frame = []
    max_eom_rows = 0
    for f, fr in enumerate(frames):
        rng = fr["rng"]
        plist = fr["patches"]
        recs = np.zeros(len(plist), abi.PATCH_DTYPE)
        pix = []  # per patch: canvas x,y, D0, D1, P0, P1 (noise-free and coded), normals
        for i, pt in enumerate(plist):
            vw = pt["view"]
            su, sv = pt["su0"] * R, pt["sv0"] * R
            yy, xx = np.nonzero(pt["mask"])
            gu, gv = xx + pt["u1"], yy + pt["v1"]
            dep = vw["depth"][gv, gu]
            if vw["mode"] == 0:
                d1 = int(dep.min())
                d0 = dep - d1
            else:
                d1 = int(dep.max())
                d0 = d1 - dep
            if max_depth is not None:
                # the encoder keeps a patch inside the range of the 8-bit geometry video (maxAllowedDepth): deeper
                # pixels are not part of the patch
                ok = d0 <= max_depth
                xx, yy, gu, gv, d0 = xx[ok], yy[ok], gu[ok], gv[ok], d0[ok]
            recs[i] = (pt["u0"], pt["v0"], pt["su0"], pt["sv0"], pt["u1"], pt["v1"], d1, vw["nrm"], vw["tan"],
                       vw["bit"], vw["mode"], pt["orient"], 1, 1, 0, su, sv)
            pix.append(dict(u=xx, v=yy, d0=d0.astype(np.int64), nrm=vw["normal"][gv, gu]))
        per_frame.append(dict(recs=recs, pix=pix))
    # EOM needs the per-frame extra-point count to size the atlas: computed in the loop below; reserve rows lazily
    H_extra_rows = 0

    def build_frame(f, Hb_total, want_arrays):
        fr, pf = frames[f], per_frame[f]
        rng = np.random.default_rng([seed, fo + f, 7])
        W, H = width, Hb_total * R
        occ_full = np.zeros((H, W), np.uint8)
        geo = np.zeros((M, H, W), np.uint16)
        att = np.zeros((M, 3, H, W), np.uint16)
        src_pos, src_col, src_nrm = [], [], []
        eom_extra = 0
        for i, pt in enumerate(fr["patches"]):
            rec, px = pf["recs"][i], pf["pix"][i]
            su, sv = pt["su0"] * R, pt["sv0"] * R
            cx, cy = _patch2canvas_arrays(pt["orient"], px["u"], px["v"], su, sv)
            cx, cy = cx + pt["u0"] * R, cy + pt["v0"] * R
            n = len(cx)
            d0 = px["d0"].copy()
            # far layer within surfaceThickness 4 (cfg/common/ctc-common.cfg:25): thicker where the surface is steep
            steep = 1.0 - np.abs(px["nrm"][:, rec["normal_axis"]])
            delta = np.minimum(4, np.floor(steep * 6 + rng.random(n) * 1.5)).astype(np.int64)
            d1v = d0 + delta
            # noise-free points -> source cloud
            if with_sources:
                for dv in (d0, d1v):
                    P = np.zeros((n, 3), np.int64)
                    P[:, rec["normal_axis"]] = (dv + rec["d1"]) if rec["projection_mode"] == 0 else np.maximum(rec["d1"] - dv, 0)
                    P[:, rec["tangent_axis"]] = px["u"] + rec["u1"]
                    P[:, rec["bitangent_axis"]] = px["v"] + rec["v1"]
                    src_pos.append(P)
                    src_nrm.append(px["nrm"])
            # coding noise
            nz = rng.random(n) < noise_fraction
            d0c = np.maximum(d0 + np.where(nz, rng.integers(-1, 2, n), 0), 0)
            nz1 = rng.random(n) < noise_fraction
            d1c = np.maximum(d1v + np.where(nz1, rng.integers(-1, 2, n), 0), d0c)
            occv = np.ones(n, np.int64)
            if eom:
                diff = d1c - d0c
                bits = np.maximum(diff - 1, 0)
                symbol = (rng.integers(0, 1 << 16, n) % (1 << bits)).astype(np.int64)
                occv = np.where(diff > 1, (1 << bits) - symbol, 1)
                eom_extra += int(sum(bin(int(s)).count("1") for s in symbol[diff > 1]))
            if want_arrays:
                occ_full[cy, cx] = occv
                geo[0][cy, cx] = d0c
                if M > 1:
                    geo[1][cy, cx] = d1c if absolute_d1 else (d1c - d0c)
                for m, dv in enumerate((d0c, d1c)[:M]):
                    P = np.zeros((n, 3), np.float64)
                    P[:, rec["normal_axis"]] = (dv + rec["d1"]) if rec["projection_mode"] == 0 else np.maximum(rec["d1"] - dv, 0)
                    P[:, rec["tangent_axis"]] = px["u"] + rec["u1"]
                    P[:, rec["bitangent_axis"]] = px["v"] + rec["v1"]
                    c16 = _yuv16_field(P, rng, color_noise)
                    for c in range(3):
                        att[m][c][cy, cx] = c16[:, c]
        out = dict(eom_extra=eom_extra)
        if want_arrays:
            # low-resolution occupancy video + padding of the dilated pixels (geometry/attribute dilation as the
            # encoder does): every pixel that the upsampled occupancy turns on must carry plausible values.
            p = p_occ
            if p > 1:
                occ_lo = occ_full.reshape(H // p, p, W // p, p).max(axis=(1, 3))
                up = np.repeat(np.repeat(occ_lo, p, axis=0), p, axis=1)
                pad = (up != 0) & (occ_full == 0)
                if pad.any():
                    # nearest valid sample inside the same p x p cell (cell-wise max works as a fill)
                    for arr in [geo[m] for m in range(M)] + [att[m][c] for m in range(M) for c in range(3)]:
                        cell = arr.reshape(H // p, p, W // p, p).max(axis=(1, 3))
                        fill = np.repeat(np.repeat(cell, p, axis=0), p, axis=1)
                        arr[pad] = fill[pad]
            else:
                occ_lo = occ_full
            out.update(occ=occ_lo, geo=geo, att=att)
        if with_sources and src_pos:
            P = np.concatenate(src_pos)
            N = np.concatenate(src_nrm)
            key = (P[:, 0] << 32) | (P[:, 1] << 16) | P[:, 2]
            _, first = np.unique(key, return_index=True)
            first = rng.permutation(first)  # PLY order shuffled by the seed
            P, N = P[first], N[first]
            out["source"] = dict(positions=P.astype(np.int16), normals=N.astype(np.float32),
                                 colors=yuv16_to_rgb8(_yuv16_field(P.astype(np.float64), None, 0)))
        return out

    if eom:
        extra = max(build_frame(f, Hb, False)["eom_extra"] for f in range(n_frames))
        H_extra_rows += -(-extra // (Wb * R * R)) + 1
    eom_v0 = Hb
    raw_v0 = Hb + H_extra_rows
    if raw_points:
        H_extra_rows += raw_rows
    Hb_total = Hb + H_extra_rows
    if height_blocks is not None:
        if height_blocks < Hb_total:
            raise ValueError(f"height_blocks {height_blocks} < needed {Hb_total}")
        Hb_total = height_blocks
    H = Hb_total * R

    occs, geos, atts, srcs = [], [], [], []
    raw_recs, raw_off = [], [0]
    eom_recs, eom_off, eom_mem = [], [0], []
    poff = [0]
    allp = []
    for f in range(n_frames):
        b = build_frame(f, Hb_total, True)
        rngx = np.random.default_rng([seed, fo + f, 11])
        if eom:
            # EOM extra points are coloured from attribute frame 0 at synthetic addresses of the EOM rectangle
            rows = slice(eom_v0 * R, (eom_v0 + max(1, H_extra_rows - (raw_rows if raw_points else 0))) * R)
            b["att"][0][:, rows, :] = rngx.integers(0, 65536, size=b["att"][0][:, rows, :].shape, dtype=np.uint16)
            npat = len(frames[f]["patches"])
            members = list(rngx.permutation(npat)) if npat else []
            eom_recs.append((0, eom_v0, len(eom_mem), len(members), b["eom_extra"]))
            eom_mem += [int(m) for m in members]
        eom_off.append(len(eom_recs))
        if raw_points:
            nraw = raw_points
            rows = slice(raw_v0 * R, (raw_v0 + raw_rows) * R)
            vals = rngx.integers(0, cube, size=3 * nraw).astype(np.uint16)
            flat = np.zeros(raw_rows * R * width, np.uint16)
            flat[:3 * nraw] = vals
            b["geo"][0][rows, :] = flat.reshape(raw_rows * R, width)
            b["att"][0][:, rows, :] = rngx.integers(0, 65536, size=b["att"][0][:, rows, :].shape, dtype=np.uint16)
            raw_recs.append((0, raw_v0, Wb, raw_rows, 0, 0, 0, nraw))
        raw_off.append(len(raw_recs))
        occs.append(b["occ"])
        geos.append(b["geo"])
        atts.append(b["att"])
        srcs.append(b.get("source"))
        allp.append(per_frame[f]["recs"])
        poff.append(poff[-1] + len(per_frame[f]["recs"]))

    params = default_params(width, H, bitdepth, p_occ)
    params.map_count_minus1 = M - 1
    params.absolute_d1 = 1 if absolute_d1 else 0
    params.patch_precedence_reverse = 1 if precedence_reverse else 0
    params.flag_color_smoothing = params.apply_attr_smoothing = 1 if color_smoothing else 0
    params.flag_geometry_smoothing = params.apply_geo_smoothing = 1 if geometry_smoothing else 0
    params.attr_transfer_filter_type = transfer_filter
    if eom:
        # lossless-style variant (cfg/common/ctc-common-lossless-geometry-attribute.cfg): no smoothing
        params.enhanced_occupancy_map_code = 1
        params.eom_fix_bit_count = 2
        params.remove_duplicate_points = 0
    if raw_points:
        params.use_additional_points_patch = 1
    gof.params = params
    gof.n_frames = n_frames
    gof.occupancy = np.ascontiguousarray(np.stack(occs))
    gof.geometry = np.ascontiguousarray(np.stack(geos))
    gof.attribute = np.ascontiguousarray(np.stack(atts))
    gof.patches = np.ascontiguousarray(np.concatenate(allp)) if allp else np.zeros(0, abi.PATCH_DTYPE)
    gof.patch_offset = np.array(poff, np.int32)
    if eom:
        gof.eom_patches = np.array(eom_recs, dtype=abi.EOM_DTYPE)
        gof.eom_offset = np.array(eom_off, np.int32)
        gof.eom_members = np.array(eom_mem if eom_mem else [0], np.int32)
    if raw_points:
        gof.raw_patches = np.array(raw_recs, dtype=abi.RAW_DTYPE)
        gof.raw_offset = np.array(raw_off, np.int32)
    gof.sources = srcs
    return gof


def concat_gofs(gofs):
    """join GOFs of identical parameters / atlas size into one (frames in list order)"""
    g0 = gofs[0]
    out = SyntheticGOF()
    out.params = g0.params
    for g in gofs[1:]:
        if bytes(g.params) != bytes(g0.params):
            raise ValueError("concat_gofs: parameter mismatch (fix the atlas height with height_blocks)")
    out.n_frames = sum(g.n_frames for g in gofs)
    out.occupancy = np.concatenate([g.occupancy for g in gofs])
    out.geometry = np.concatenate([g.geometry for g in gofs])
    out.attribute = np.concatenate([g.attribute for g in gofs])
    out.patches = np.concatenate([g.patches for g in gofs])

    def cat_off(name):
        offs, base = [0], 0
        for g in gofs:
            o = getattr(g, name)
            offs += [int(x) + base for x in o[1:]]
            base += int(o[-1])
        return np.array(offs, np.int32)
    out.patch_offset = cat_off("patch_offset")
    if g0.eom_patches is not None:
        recs, mem, base = [], [], 0
        for g in gofs:
            r = g.eom_patches.copy()
            r["member_begin"] += base
            base += len(g.eom_members)
            recs.append(r)
            mem.append(g.eom_members)
        out.eom_patches = np.concatenate(recs)
        out.eom_members = np.concatenate(mem)
        out.eom_offset = cat_off("eom_offset")
    if g0.raw_patches is not None:
        out.raw_patches = np.concatenate([g.raw_patches for g in gofs])
        out.raw_offset = cat_off("raw_offset")
    out.sources = [s for g in gofs for s in g.sources]
    return out


def slice_gof(g, begin, end):
    """frames [begin, end) of a GOF as a new GOF (views, no plane copies)"""
    out = SyntheticGOF()
    out.params = g.params
    out.n_frames = end - begin
    out.occupancy, out.geometry, out.attribute = g.occupancy[begin:end], g.geometry[begin:end], g.attribute[begin:end]

    def cut(off, recs):
        o = off[begin:end + 1]
        return np.ascontiguousarray(recs[o[0]:o[-1]]), (o - o[0]).astype(np.int32)
    out.patches, out.patch_offset = cut(g.patch_offset, g.patches)
    if g.eom_patches is not None:
        out.eom_patches, out.eom_offset = cut(g.eom_offset, g.eom_patches)
        out.eom_members = g.eom_members
    if g.raw_patches is not None:
        out.raw_patches, out.raw_offset = cut(g.raw_offset, g.raw_patches)
    out.sources = g.sources[begin:end]
    return out


def _gen_one(kw):
    return generate_gof(**kw)


def generate_gof_parallel(n_frames, workers=None, **kw):
    """generate_gof frame by frame in worker processes (same result as one call with a fixed `height_blocks`)"""
    import multiprocessing as mp
    import os
    if "height_blocks" not in kw or kw["height_blocks"] is None:
        raise ValueError("generate_gof_parallel needs a fixed height_blocks")
    workers = workers or min(n_frames, os.cpu_count() or 1)
    jobs = [dict(kw, n_frames=1, frame_offset=kw.get("frame_offset", 0) + f) for f in range(n_frames)]
    if workers <= 1:
        parts = [_gen_one(j) for j in jobs]
    else:
        with mp.get_context("fork").Pool(workers) as pool:
            parts = pool.map(_gen_one, jobs)
    return concat_gofs(parts)


def to_decoder_planes(gof, bitdepth=8, filt=0, sample_dtype=np.uint8):
    """The GOF as a video decoder would hand it over (before PCCVideoDecoder's inverse colour conversion): geometry luma
    and 4:2:0 attribute frames { Y [H][W], U [H/2][W/2], V [H/2][W/2] } of `bitdepth`-bit samples, made from the GOF's
    16-bit 4:4:4 attribute frames by dropping the low bits and every second chroma row / column.  Returns the native
    planes for PCCCodecB200.uploadGofYuv420; the GOF's own 4:4:4 frames must then be REPLACED by the conversion of these
    planes (oracle / reference) before it is used as the expected input."""
    F, M = gof.n_frames, gof.params.map_count_minus1 + 1
    H, W = gof.params.height, gof.params.width
    out = {"bitdepth": bitdepth, "filter": filt, "geometry": None, "attribute": None}
    g = gof.geometry.reshape(F, M, H, W)
    out["geometry"] = np.ascontiguousarray(g.astype(sample_dtype)) if g.max() < (1 << (8 * np.dtype(sample_dtype).itemsize)) else np.ascontiguousarray(g)
    if gof.attribute is not None:
        a = gof.attribute.reshape(F, M, 3, H, W)
        sh = 16 - bitdepth
        frames = np.empty((F, M, H * W + 2 * (H // 2) * (W // 2)), sample_dtype)
        frames[:, :, :H * W] = (a[:, :, 0] >> sh).reshape(F, M, -1)
        q = (H // 2) * (W // 2)
        frames[:, :, H * W:H * W + q] = (a[:, :, 1, ::2, ::2] >> sh).reshape(F, M, -1)
        frames[:, :, H * W + q:] = (a[:, :, 2, ::2, ::2] >> sh).reshape(F, M, -1)
        out["attribute"] = np.ascontiguousarray(frames)
    return out


def make_relative_t1(gof, seed=0):
    """Turns the GOF into the multiple-streams layout with a delta-coded second attribute map (CTC condition
    T1-from-rec-T0; colorPointCloud, PCCCodec.cpp:1387-1416): map 1 becomes clip(T1 - T0 + 32768), with a sprinkling of
    extreme codes so that both clamps of the reconstruction are exercised."""
    p = gof.params
    F, M, H, W = gof.n_frames, p.map_count_minus1 + 1, p.height, p.width
    assert M == 2 and gof.attribute is not None
    a = gof.attribute.reshape(F, M, 3, H, W)
    rng = np.random.default_rng([seed, 991])
    delta = a[:, 1].astype(np.int64) - a[:, 0].astype(np.int64) + 32768
    code = np.clip(delta, 0, 65535)
    r = rng.random(code.shape)
    code = np.where(r < 0.003, 0, np.where(r > 0.997, 65535, code))
    a[:, 1] = code.astype(np.uint16)
    gof.attribute = np.ascontiguousarray(a.reshape(gof.attribute.shape))
    p.multiple_streams = 1
    p.relative_t1 = 1
    return gof


def make_pixel_interleaved(gof, surface_thickness=4):
    """Turns a two-map GOF into the singleMapPixelInterleaving layout (generatePoints, PCCCodec.cpp:350-471): one map
    whose pixels carry the near layer where (x + y) is even and the far layer where it is odd."""
    p = gof.params
    F, M, H, W = gof.n_frames, p.map_count_minus1 + 1, p.height, p.width
    assert M == 2
    odd = ((np.arange(H)[:, None] + np.arange(W)[None, :]) & 1).astype(bool)
    g = gof.geometry.reshape(F, M, H, W)
    gof.geometry = np.ascontiguousarray(np.where(odd, g[:, 1], g[:, 0]).reshape(F, 1, H, W))
    if gof.attribute is not None:
        a = gof.attribute.reshape(F, M, 3, H, W)
        gof.attribute = np.ascontiguousarray(np.where(odd, a[:, 1], a[:, 0]).reshape(F, 1, 3, H, W))
    p.map_count_minus1 = 0
    p.single_map_pixel_interleaving = 1
    p.surface_thickness = surface_thickness
    return gof


def make_plr(gof, seed=0):
    """Turns a single-map GOF into a point-local-reconstruction one (generatePoints, PCCCodec.cpp:472-496): the mode
    table of PCCDecoder::setPointLocalReconstruction (entry 0 = no second point) and a random mode per patch block."""
    p = gof.params
    assert p.map_count_minus1 == 0
    rng = np.random.default_rng([seed, 4471])
    # (interpolate, filling, minD1, neighbor)
    gof_modes = np.array([[0, 0, 0, 1], [1, 0, 0, 1], [1, 1, 0, 1], [1, 1, 1, 2], [0, 1, 2, 1], [1, 0, 3, 2], [0, 0, 1, 1]], np.uint8)
    nblk = gof.patches["size_u0"].astype(np.int64) * gof.patches["size_v0"].astype(np.int64)
    off = np.zeros(len(nblk) + 1, np.int64)
    off[1:] = np.cumsum(nblk)
    bm = rng.integers(0, len(gof_modes), int(off[-1])).astype(np.uint8)
    for i in np.nonzero(rng.random(len(nblk)) < 0.3)[0]:  # patch-level mode (level flag 1): every block the same
        bm[off[i]:off[i + 1]] = rng.integers(0, len(gof_modes))
    # getDeltaNeighbors (:249-252) reads one row past the frame for pixels within `neighbor` of the bottom edge (its
    # bounds are inclusive): undefined in the reference, so patches that reach the last block row do not interpolate
    Hb = p.height // p.occupancy_resolution
    swapped = np.isin(gof.patches["orientation"], (1, 2, 4, 6, 8))
    rows = np.where(swapped, gof.patches["size_u0"], gof.patches["size_v0"])
    for i in np.nonzero(gof.patches["v0"] + rows >= Hb)[0]:
        blk = bm[off[i]:off[i + 1]]
        blk[gof_modes[blk, 0] != 0] = 4
    gof.plr = dict(modes=gof_modes, block_mode=bm, block_offset=off)
    p.point_local_reconstruction = 1
    return gof


def make_aux_video(gof, seed=0):
    """Moves the raw patches (and the EOM patches' colours) of a GOF generated with raw_points / eom into an auxiliary video
    (asps.getAuxiliaryVideoEnabledFlag: PCCCodec.cpp:895-897 reads the raw coordinates from
    context.getVideoRawPointsGeometry(), :1524-1580 the raw and EOM colours from the auxiliary attribute video through
    8-bit PCCColor3B values).  The auxiliary frames are 64-aligned as the syntax requires (auxiliaryVideoTileRowWidthMinus1 /
    RowHeight in units of 64, PCCDecoder.cpp:1820-1824); the attribute samples carry high bits so that the truncation shows."""
    have_raw = gof.raw_patches is not None and len(gof.raw_patches) > 0
    have_eom = gof.eom_patches is not None and len(gof.eom_patches) > 0
    assert have_raw or have_eom
    p, R = gof.params, gof.params.occupancy_resolution
    rng = np.random.default_rng(seed + 4242)
    F, M, W = gof.n_frames, p.map_count_minus1 + 1, p.width
    geo = gof.geometry.reshape(F, M, p.height, W)
    Wa = -(-W // 64) * 64
    raw_rows = int(max(r["size_v0"] for r in gof.raw_patches)) if have_raw else 0     # in blocks; raw patch at block row 1
    eom_v0 = 1 + raw_rows + 1                                                          # EOM patch below it
    eom_rows = -(-int(max(e["eom_count"] for e in gof.eom_patches)) // (R * R * (Wa // R))) + 1 if have_eom else 0
    Ha = -(-((eom_v0 + eom_rows + 1) * R) // 64) * 64
    gof.aux_geometry = rng.integers(0, 1 << 10, size=(F, Ha, Wa)).astype(np.uint16) if have_raw else None
    gof.aux_attribute = rng.integers(0, 1 << 16, size=(F, 3, Ha, Wa)).astype(np.uint16)
    for f in range(F):
        for k in range(int(gof.raw_offset[f]), int(gof.raw_offset[f + 1])) if have_raw else ():
            r = gof.raw_patches[k]
            rows = int(r["size_v0"]) * R
            src = geo[f, 0, int(r["v0"]) * R:int(r["v0"]) * R + rows, int(r["u0"]) * R:(int(r["u0"]) + int(r["size_u0"])) * R]
            gof.aux_geometry[f, R:R + rows, :src.shape[1]] = src
            gof.raw_patches[k]["u0"], gof.raw_patches[k]["v0"] = 0, 1
        for k in range(int(gof.eom_offset[f]), int(gof.eom_offset[f + 1])) if have_eom else ():
            gof.eom_patches[k]["u0"], gof.eom_patches[k]["v0"] = 1, eom_v0
    p.use_aux_separate_video, p.aux_width, p.aux_height = 1, Wa, Ha
    return gof


def make_lod_and_oblique(gof, seed=0, lod=True, planes=True):
    """Turns some patches of a generated GOF into level-of-detail-2 patches (PCCPatch::generatePoint, PCCPatch.h:201-207:
    u * lodX + u1, v * lodY + v1) and / or patches of the 45-degree projection planes (axisOfAdditionalPlane 1..3:
    PCCCodec::inverseRotatePosition45DegreeOnAxis, PCCCodec.cpp:2503-2524, with its wrap of negative intermediates).
    The planes stay as they are: the patches simply decode to other positions — what matters is that the CUDA path and
    the reference decode the same thing.  A patch only gets lod 2 when its points stay inside the cube."""
    rng = np.random.default_rng([seed, 0x10D])
    cube = 1 << gof.params.geometry_bitdepth_3d
    R = gof.params.occupancy_resolution
    for p in gof.patches:
        if lod:
            if p["u1"] + 2 * p["size_u0"] * R < cube and rng.random() < 0.5:
                p["lod_x"] = 2
            if p["v1"] + 2 * p["size_v0"] * R < cube and rng.random() < 0.5:
                p["lod_y"] = 2
        if planes and rng.random() < 0.6:
            p["axis_of_additional_plane"] = int(rng.integers(1, 4))
    return gof


def generate_stacked_gof(n_patches=12, n_frames=1, seed=0, luma=65535, transfer_filter=0):
    """A pathological GOF: `n_patches` patches of 32x32 pixels that all decode to (nearly) the same 3-D slab, with a
    saturated luma.  Every 4^3 colour cell then holds a few hundred points whose float luma sum exceeds 2^24, the range
    in which the reference's float accumulators (PCCCodec.cpp:996, :1177) are exact — the case in which the order of the
    additions shows in the result."""
    R, W, H, bitdepth = 16, 256, 256, 8
    p = default_params(W, H, bitdepth, 4)
    p.attr_transfer_filter_type = transfer_filter
    M = 2
    g = SyntheticGOF()
    g.params, g.n_frames = p, n_frames
    g.occupancy = np.zeros((n_frames, H // 4, W // 4), np.uint8)
    g.geometry = np.zeros((n_frames, M, H, W), np.uint16)
    g.attribute = np.zeros((n_frames, M, 3, H, W), np.uint16)
    recs = np.zeros(n_frames * n_patches, abi.PATCH_DTYPE)
    rng = np.random.default_rng([seed, 0x57AC])
    for f in range(n_frames):
        for k in range(n_patches):
            bx, by = (k % 8) * 2, (k // 8) * 2
            nrm, tan, bit, mode = VIEW_AXES[0]
            # (u0, v0, sizeU0, sizeV0, u1, v1, d1, axes, mode, orientation, lod, lod, plane, size2d)
            recs[f * n_patches + k] = (bx, by, 2, 2, 64 + (k % 3), 64, 40 + (k % 2), nrm, tan, bit, mode, 0, 1, 1, 0, 32, 32)
            ys, xs = slice(by * R, by * R + 32), slice(bx * R, bx * R + 32)
            g.occupancy[f, by * 4:by * 4 + 8, bx * 4:bx * 4 + 8] = 1
            d0 = (4 + rng.integers(0, 3, size=(32, 32))).astype(np.uint16)
            g.geometry[f, 0, ys, xs] = d0
            g.geometry[f, 1, ys, xs] = d0 + rng.integers(0, 3, size=(32, 32)).astype(np.uint16)
            for m in range(M):
                g.attribute[f, m, 0, ys, xs] = np.clip(luma - 257 * k - rng.integers(0, 600, size=(32, 32)), 0, 65535)
                g.attribute[f, m, 1, ys, xs] = 30000 + rng.integers(0, 2000, size=(32, 32))
                g.attribute[f, m, 2, ys, xs] = 34000 + rng.integers(0, 2000, size=(32, 32))
    g.patches = recs
    g.patch_offset = np.arange(0, (n_frames + 1) * n_patches, n_patches, dtype=np.int32)
    g.sources = []
    return g
