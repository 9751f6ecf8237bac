"""ctypes mirror of include/mitransient_b200.h (data definitions only, no behaviour)."""
from __future__ import annotations

import ctypes as C

TR_ABI_VERSION = 14

TR_OK = 0
TR_ERR_INVALID_ARGUMENT = -1
TR_ERR_UNSUPPORTED = -2
TR_ERR_CUDA = -3
TR_ERR_NO_DEVICE = -4
TR_ERR_OUT_OF_MEMORY = -5

TR_SHAPE_MESH, TR_SHAPE_RECTANGLE = 0, 1
TR_PRIM_TRIANGLE, TR_PRIM_QUAD, TR_PRIM_BOX = 0, 1, 2
TR_SHAPE_FLAG_SENSOR = 1
TR_BSDF_DIFFUSE, TR_BSDF_NULL, TR_BSDF_CONDUCTOR, TR_BSDF_DIELECTRIC = 0, 1, 2, 3
TR_BSDF_FLAG_TWOSIDED = 1
TR_EMITTER_AREA, TR_EMITTER_PROJECTOR, TR_EMITTER_ANGULARAREA = 0, 1, 2
TR_SENSOR_PERSPECTIVE, TR_SENSOR_NLOS_CAPTURE_METER = 0, 1
TR_INTEGRATOR_TRANSIENT_PATH, TR_INTEGRATOR_TRANSIENT_NLOS_PATH, TR_INTEGRATOR_TRANSIENT_PRBVOLPATH = 0, 1, 2

TR_FLAG_CAMERA_UNWARP = 1
TR_FLAG_DISCARD_DIRECT_LIGHT = 2
TR_FLAG_HIDE_EMITTERS = 4
TR_FLAG_NLOS_LASER_SAMPLING = 8
TR_FLAG_NLOS_HG_SAMPLING = 16
TR_FLAG_NLOS_HG_SAMPLING_DO_RROULETTE = 32
TR_FLAG_NLOS_HG_SAMPLING_INCLUDES_RELAY_WALL = 64
TR_FLAG_ACCOUNT_FIRST_AND_LAST_BOUNCES = 128
TR_FLAG_DISCARD_DIRECT_PATHS = 256
TR_HIST_AUTO, TR_HIST_SHARED, TR_HIST_GLOBAL = 0, 1, 2
HIST_NAMES = {"auto": TR_HIST_AUTO, "shared": TR_HIST_SHARED, "global": TR_HIST_GLOBAL}

TR_CAPTURE_SINGLE, TR_CAPTURE_CONFOCAL, TR_CAPTURE_EXHAUSTIVE = 0, 1, 2
TR_FILM_TRANSIENT_HDR, TR_FILM_PHASOR_HDR = 0, 1


class TrShapeDesc(C.Structure):
    _fields_ = [("type", C.c_int32), ("bsdf", C.c_int32), ("emitter", C.c_int32), ("flags", C.c_int32),
                ("interior_medium", C.c_int32), ("exterior_medium", C.c_int32),
                ("prim_begin", C.c_uint32), ("prim_count", C.c_uint32), ("to_world", C.c_float * 12)]


class TrBsdfDesc(C.Structure):
    _fields_ = [("type", C.c_int32), ("flags", C.c_int32), ("reflectance", C.c_float * 3), ("eta", C.c_float * 3),
                ("k", C.c_float * 3), ("specular_transmittance", C.c_float * 3), ("_pad", C.c_float * 2)]


class TrEmitterDesc(C.Structure):
    _fields_ = [("type", C.c_int32), ("shape", C.c_int32), ("radiance", C.c_float * 3), ("fov_x", C.c_float),
                ("to_world", C.c_float * 12), ("world_to_local", C.c_float * 12), ("camera_to_sample", C.c_float * 16),
                ("cutoff_angle", C.c_float), ("beam_width", C.c_float), ("cos_cutoff_angle", C.c_float),
                ("cos_beam_width", C.c_float), ("inv_transition_width", C.c_float), ("_pad", C.c_float)]


class TrMediumDesc(C.Structure):
    _fields_ = [("sigma_t", C.c_float * 3), ("albedo", C.c_float * 3), ("g", C.c_float), ("scale", C.c_float)]


class TrSensorDesc(C.Structure):
    _fields_ = [("type", C.c_int32), ("shape", C.c_int32), ("to_world", C.c_float * 12),
                ("sample_to_camera", C.c_float * 16), ("near_clip", C.c_float), ("far_clip", C.c_float),
                ("origin", C.c_float * 3), ("medium", C.c_int32), ("is_confocal", C.c_int32), ("laser_target", C.c_float * 3)]


class TrFilmDesc(C.Structure):
    _fields_ = [("width", C.c_uint32), ("height", C.c_uint32), ("temporal_bins", C.c_uint32),
                ("start_opl", C.c_float), ("bin_width_opl", C.c_float),
                ("laser_scan_width", C.c_uint32), ("laser_scan_height", C.c_uint32),
                ("type", C.c_uint32), ("n_frequencies", C.c_uint32),
                ("crop_offset_x", C.c_uint32), ("crop_offset_y", C.c_uint32), ("_pad", C.c_uint32),
                ("frequencies", C.POINTER(C.c_float))]


class TrSceneDesc(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("n_prims", C.c_uint32), ("prim_verts", C.POINTER(C.c_float)),
                ("prim_types", C.POINTER(C.c_uint8)),
                ("n_shapes", C.c_uint32), ("shapes", C.POINTER(TrShapeDesc)),
                ("n_bsdfs", C.c_uint32), ("bsdfs", C.POINTER(TrBsdfDesc)),
                ("n_emitters", C.c_uint32), ("emitters", C.POINTER(TrEmitterDesc)),
                ("n_media", C.c_uint32), ("media", C.POINTER(TrMediumDesc)),
                ("sensor", TrSensorDesc), ("film", TrFilmDesc)]


class TrRenderParams(C.Structure):
    _fields_ = [("integrator", C.c_int32), ("max_depth", C.c_uint32), ("rr_depth", C.c_uint32),
                ("flags", C.c_uint32), ("spp", C.c_uint32), ("total_spp", C.c_uint32), ("seed", C.c_uint32),
                ("tile_size", C.c_uint32), ("rank", C.c_uint32), ("world", C.c_uint32),
                ("film_channels", C.c_int32), ("accumulate", C.c_int32),
                ("capture_type", C.c_int32), ("filter_depth", C.c_int32), ("force_equal_grids", C.c_int32),
                ("laser_scan_sample_to_camera", C.c_float * 16), ("histogram_path", C.c_uint32), ("_reserved", C.c_uint32 * 5)]


class TrRenderInfo(C.Structure):
    _fields_ = [("histogram_path", C.c_uint32), ("slots", C.c_uint32), ("threads", C.c_uint32), ("smem_bytes", C.c_uint32)]


class TrCounters(C.Structure):
    _fields_ = [("paths", C.c_uint64), ("vertices", C.c_uint64), ("deposits", C.c_uint64),
                ("deposits_in_range", C.c_uint64), ("bin_checksum", C.c_uint64), ("shadow_rays", C.c_uint64),
                ("reserved", C.c_uint64 * 2)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k in
                ("paths", "vertices", "deposits", "deposits_in_range", "bin_checksum", "shadow_rays")}


class TrDeposit(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("opl", C.c_float),
                ("r", C.c_float), ("g", C.c_float), ("b", C.c_float)]


# every symbol include/mitransient_b200.h declares (checked by tests/test_abi.py)
EXPORTED_SYMBOLS = (
    "tr_device_check", "tr_abi_version", "tr_last_error", "tr_scene_create", "tr_scene_destroy",
    "tr_scene_info", "tr_render", "tr_render_to_host", "tr_owned_pixels", "tr_film_scatter_add", "tr_film_develop",
    "tr_peer_alloc", "tr_peer_open", "tr_peer_close", "tr_peer_free", "tr_selftest_rcp", "tr_scene_refresh", "tr_last_render_info",
)
TR_PEER_HANDLE_BYTES = 64
