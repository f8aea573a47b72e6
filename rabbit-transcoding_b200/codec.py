"""Host-side mirror of the reference's PCCCodec entry points for the hot path, over the C ABI.

The method names and argument meaning follow the reference (PccLibCommon/include/PCCCodec.h:160-183 and the
decoder's frame loop, PccLibDecoder/source/PCCDecoder.cpp:330-508) so the parity tests read like the
reference's own call sequence:

    codec = PCCCodecB200(device=0)
    codec.beginGof(params, n_frames); codec.uploadFrames(frames, atlas)
    codec.generatePointCloud()            # generateOccupancyMap + generateBlockToPatch... + generatePointCloud
                                          #   + colorPointCloud for every frame of the GOF
    codec.smoothPointCloudPostprocess()   # PCCCodec.cpp:52-147
    codec.transferColors16bitBP()         # PCCPointSet.cpp:1126-1485 as called at PCCDecoder.cpp:447-465
    codec.colorSmoothing()                # PCCCodec.cpp:149-236
    codec.convertYUV16ToRGB8()            # PCCPointSet.h:133-166
    cloud = codec.getPointCloud(f)        # dict of numpy arrays in PCCPointSet3 layouts

Error behaviour: the reference prints and exits; here every non-zero status raises `RabbitError` carrying the
status code (180 for a patch outside the canvas, as PCCPatch.cpp:237-245 exits with).
There is no CPU fallback: constructing a codec without the CUDA library or without a GPU raises.
"""
import ctypes as C

import numpy as np

from . import abi


class RabbitError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"rabbit_b200 status {status}: {message}")
        self.status = status


class PCCCodecB200:
    def __init__(self, device=0, stream=None):
        self._lib = abi.load_library()
        if self._lib.rb200_abi_version() != 3:
            raise RuntimeError("rabbit_b200 ABI version mismatch")
        h = C.c_void_p()
        st = self._lib.rb200_create(device, C.byref(h))
        if st != abi.RB200_OK:
            raise RabbitError(st, "rb200_create failed (no CUDA device? there is no CPU fallback)")
        self._h = h
        self.device = device
        self.n_frames = 0
        self.params = None
        self._keep = None
        if stream is not None:
            self.setStream(stream)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.rb200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        if st != abi.RB200_OK:
            raise RabbitError(st, self._lib.rb200_error_string(self._h).decode())

    # ---- plumbing ----
    def setStream(self, cuda_stream):
        self._check(self._lib.rb200_set_stream(self._h, C.c_void_p(int(cuda_stream) if cuda_stream else 0)))

    def synchronize(self):
        self._check(self._lib.rb200_synchronize(self._h))

    def beginGof(self, params, n_frames):
        self._check(self._lib.rb200_gof_begin(self._h, C.byref(params), n_frames))
        self.params, self.n_frames = params, n_frames

    def uploadFrames(self, frames, atlas, keep=None):
        """frames: abi.Frames (host or device pointers), atlas: abi.Atlas (host pointers)"""
        self._keep = keep
        self._check(self._lib.rb200_gof_upload(self._h, C.byref(frames), C.byref(atlas)))

    def uploadGof(self, gof):
        self.beginGof(gof.params, gof.n_frames)
        self.uploadFrames(gof.frames_struct(), gof.atlas_struct(), keep=gof)
        if getattr(gof, "plr", None) is not None:
            self.setPointLocalReconstruction(gof.plr)

    def setPointLocalReconstruction(self, plr):
        """PCCDecoder::setPointLocalReconstruction / setPLRData (PCCDecoder.cpp:528-591): the mode table and the mode of
        every patch block; plr = dict(modes=uint8[n, 4] (interpolate, filling, minD1, neighbor), block_mode=uint8[],
        block_offset=int64[patches + 1])"""
        self._check(self._lib.rb200_gof_set_plr(self._h, C.byref(abi.plr_struct(plr))))

    def uploadGofYuv420(self, gof, native):
        """decoder-native planes (synthetic.to_decoder_planes): 4:2:0 attribute frames of 8/10-bit samples and geometry
        luma; PCCVideoDecoder's inverse colour conversion (PCCVideoDecoder.cpp:125-146) runs on the device"""
        self.beginGof(gof.params, gof.n_frames)
        f = abi.FramesYuv420()
        f.occupancy = abi.ptr(gof.occupancy)
        f.geometry = abi.ptr(native["geometry"])
        f.attribute = abi.ptr(native["attribute"]) if native.get("attribute") is not None else None
        f.geometry_sample_bytes = native["geometry"].dtype.itemsize
        f.attribute_sample_bytes = native["attribute"].dtype.itemsize if native.get("attribute") is not None else 1
        f.attribute_bitdepth = native["bitdepth"]
        f.upsampling_filter = native["filter"]
        f.geometry_shift = native.get("geometry_shift", 0)
        f.attribute_shift = native.get("attribute_shift", 0)
        # PCCImage::convertBitdepth of the geometry / occupancy video (PCCDecoder.cpp:119, :148-149): (in, out, msbAlign)
        gb, ob = native.get("geometry_bitdepth"), native.get("occupancy_bitdepth")
        if gb is not None:
            f.geometry_bitdepth_in, f.geometry_bitdepth_out, f.geometry_msb_align = gb
        if ob is not None:
            f.occupancy_bitdepth_out, f.occupancy_msb_align = ob
        self._keep = (gof, native)
        self._check(self._lib.rb200_gof_upload_yuv420(self._h, C.byref(f), C.byref(gof.atlas_struct())))

    def uploadGofNv12(self, gof, surfaces):
        """pitched NV12 / P010 decoder surfaces in device (or pinned) memory — what NVDEC leaves behind in RABBIT's --useCuda
        path: surfaces = dict(occupancy=[F x (luma_ptr, pitch)], geometry=[F*M x (luma_ptr, pitch)],
        attribute=[F*M x (luma_ptr, pitch, chroma_ptr, pitch)], sample_bytes, sample_lsb_shift, bitdepth, filter, ...)"""
        self.beginGof(gof.params, gof.n_frames)

        def table(rows):
            arr = (abi.Surface * len(rows))()
            for i, r in enumerate(rows):
                arr[i].luma, arr[i].pitch_luma = r[0], r[1]
                if len(r) > 2:
                    arr[i].chroma, arr[i].pitch_chroma = r[2], r[3]
            return arr
        occ, geo = table(surfaces["occupancy"]), table(surfaces["geometry"])
        att = table(surfaces["attribute"]) if surfaces.get("attribute") else None
        f = abi.FramesNv12()
        f.occupancy, f.geometry = C.cast(occ, C.c_void_p), C.cast(geo, C.c_void_p)
        f.attribute = C.cast(att, C.c_void_p) if att is not None else None
        f.sample_bytes = surfaces.get("sample_bytes", 1)
        f.sample_lsb_shift = surfaces.get("sample_lsb_shift", 0)
        f.conversion.attribute_bitdepth = surfaces.get("bitdepth", 8)
        f.conversion.upsampling_filter = surfaces.get("filter", 0)
        f.conversion.geometry_shift = surfaces.get("geometry_shift", 0)
        f.conversion.attribute_shift = surfaces.get("attribute_shift", 0)
        self._keep = (gof, surfaces, occ, geo, att)
        self._check(self._lib.rb200_gof_upload_nv12(self._h, C.byref(f), C.byref(gof.atlas_struct())))

    def getPlanes(self, frame, m):
        """geometry [H][W] and attribute [3][H][W] uint16 as they sit in HBM after the upload"""
        p = self.params
        geo = np.zeros((p.height, p.width), np.uint16)
        att = np.zeros((3, p.height, p.width), np.uint16) if p.attribute_count > 0 else None
        self._check(self._lib.rb200_download_planes(self._h, frame, m, abi.ptr(geo), abi.ptr(att)))
        return geo, att

    # ---- the reference's entry points ----
    def generatePointCloud(self):
        self._check(self._lib.rb200_reconstruct(self._h))

    def smoothPointCloudPostprocess(self):
        self._check(self._lib.rb200_smooth_geometry(self._h))

    def transferColors16bitBP(self):
        self._check(self._lib.rb200_transfer_colors(self._h))

    def colorSmoothing(self):
        self._check(self._lib.rb200_smooth_color(self._h))

    def convertYUV16ToRGB8(self):
        self._check(self._lib.rb200_convert_rgb8(self._h))

    def decodeGof(self):
        """whole per-frame sequence of PCCDecoder.cpp:330-508 for the resident GOF"""
        self._check(self._lib.rb200_decode_gof(self._h))

    # ---- results ----
    def frameCounts(self):
        arr = (abi.FrameCounts * self.n_frames)()
        self._check(self._lib.rb200_frame_counts_get(self._h, arr))
        return list(arr)

    def getPointCloud(self, f, counts=None, fields=("positions", "colors16", "colors", "boundary_types", "partition",
                                                     "point_to_pixel")):
        counts = counts or self.frameCounts()
        n = counts[f].total
        shapes = dict(positions=((n, 3), np.int16), colors16=((n, 3), np.uint16), colors=((n, 3), np.uint8),
                      boundary_types=((n,), np.uint16), partition=((n,), np.uint32), point_to_pixel=((n, 3), np.uint32))
        out = {k: np.zeros(*shapes[k]) for k in fields}
        h = abi.CloudHost(*[abi.ptr(out[k]) if k in out else None for k, _ in abi.CloudHost._fields_])
        self._check(self._lib.rb200_download_frame(self._h, f, C.byref(h)))
        return out

    def getGof(self, counts=None, fields=("positions", "colors"), out=None):
        """all frames back to back (one packed D2H copy per field); `out` may hold preallocated (pinned) arrays"""
        counts = counts or self.frameCounts()
        n = sum(c.total for c in counts)
        shapes = dict(positions=(3, np.int16), colors16=(3, np.uint16), colors=(3, np.uint8),
                      boundary_types=(0, np.uint16), partition=(0, np.uint32), point_to_pixel=(3, np.uint32))
        res = {}
        for k in fields:
            w, dt = shapes[k]
            if out is not None and k in out:
                buf = out[k]
                if buf.shape[0] < n:
                    raise ValueError(f"getGof: out[{k!r}] holds {buf.shape[0]} points, {n} needed")
                res[k] = buf
            else:
                res[k] = np.zeros((n, w) if w else (n,), dt)
        h = abi.CloudHost(*[abi.ptr(res[k]) if k in res else None for k, _ in abi.CloudHost._fields_])
        self._check(self._lib.rb200_download_gof(self._h, C.byref(h)))
        return res, n

    def getBlockToPatch(self, f):
        p = self.params
        a = np.zeros((p.height // p.occupancy_resolution, p.width // p.occupancy_resolution), np.uint32)
        self._check(self._lib.rb200_download_block_to_patch(self._h, f, abi.ptr(a)))
        return a

    def getOccupancyMap(self, f):
        p = self.params
        a = np.zeros((p.height, p.width), np.uint8)
        self._check(self._lib.rb200_download_occupancy(self._h, f, abi.ptr(a)))
        return a

    def computeChecksum(self, f, reorderPoints=False):
        """PCCPointSet3::computeChecksum( reorderPoints ): MD5 of positions || RGB8 of frame f as a hex string; with
        reorderPoints the canonical order (sorted, duplicates merged) is hashed"""
        d = (C.c_uint8 * 16)()
        self._check((self._lib.rb200_frame_md5_canonical if reorderPoints else self._lib.rb200_frame_md5)(self._h, f, d))
        return bytes(d).hex()

    def write(self, f, file_name):
        """PCCPointSet3::write( fileName, asAscii=false ) for decoded frame f"""
        self._check(self._lib.rb200_write_ply(self._h, f, file_name.encode()))

    def read(self, file_name):
        """PCCPointSet3::read: returns dict(positions, colors or None)"""
        n, hc = abi.i64(0), C.c_int(0)
        st = self._lib.rb200_read_ply(file_name.encode(), None, None, 0, C.byref(n), C.byref(hc))
        if st != abi.RB200_OK:
            raise RabbitError(st, f"cannot read {file_name}")
        pos = np.zeros((n.value, 3), np.int16)
        col = np.zeros((n.value, 3), np.uint8) if hc.value else None
        st = self._lib.rb200_read_ply(file_name.encode(), abi.ptr(pos), abi.ptr(col), n.value, C.byref(n), C.byref(hc))
        if st != abi.RB200_OK:
            raise RabbitError(st, f"cannot read {file_name}")
        return dict(positions=pos, colors=col)

    # ---- instrumentation ----
    def stats(self, reset=False):
        s = abi.LaunchStats()
        self._check(self._lib.rb200_stats_get(self._h, C.byref(s), 1 if reset else 0))
        return s

    def enableTiming(self, on=True):
        self._check(self._lib.rb200_timing_enable(self._h, 1 if on else 0))

    def timings(self):
        out = {}
        i = 0
        name = C.create_string_buffer(64)
        ms = C.c_double()
        n = abi.i64()
        while self._lib.rb200_timing_get(self._h, i, name, 64, C.byref(ms), C.byref(n)) == abi.RB200_OK:
            out[name.value.decode()] = (ms.value, n.value)
            i += 1
        return out


class Decoder:
    """Convenience wrapper: one call = the decoder's reconstruction + post-processing of a GOF."""

    def __init__(self, device=0):
        self.codec = PCCCodecB200(device)

    def decode_gof(self, gof, fields=("positions", "colors16", "colors", "boundary_types", "partition",
                                      "point_to_pixel")):
        self.codec.uploadGof(gof)
        self.codec.decodeGof()
        counts = self.codec.frameCounts()
        return [self.codec.getPointCloud(f, counts, fields) for f in range(gof.n_frames)]
