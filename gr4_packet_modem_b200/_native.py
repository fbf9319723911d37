"""ctypes binding of include/b200sync.h.  Fails loudly when libb200sync.so is missing."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# B200SYNC_LIB: development override used by scripts/variants.sh to time alternative builds of the same ABI
LIB_PATH = os.environ.get("B200SYNC_LIB") or os.path.join(_HERE, "libb200sync.so")


class SdConfig(C.Structure):
    _fields_ = [
        ("fft_size", C.c_uint32),
        ("samples_per_symbol", C.c_uint32),
        ("rrc_taps", C.c_void_p),
        ("n_rrc_taps", C.c_uint32),
        ("syncword", C.c_void_p),
        ("n_syncword", C.c_uint32),
        ("constellation", C.c_void_p),
        ("n_constellation", C.c_uint32),
        ("min_freq_bin", C.c_int32),
        ("max_freq_bin", C.c_int32),
        ("time_threshold", C.c_uint64),
        ("power_threshold", C.c_float),
        ("device", C.c_int32),
    ]


class DetectionRecord(C.Structure):
    _fields_ = [
        ("index", C.c_uint64),
        ("corr_re", C.c_float),
        ("corr_im", C.c_float),
        ("pow", C.c_float),
        ("pow_left", C.c_float),
        ("pow_right", C.c_float),
        ("pow_prev", C.c_float),
        ("pow_next", C.c_float),
        ("noise_power", C.c_float),
        ("freq_bin", C.c_int32),
        ("_pad", C.c_int32),
    ]


class SyncwordTag(C.Structure):
    _fields_ = [
        ("index", C.c_uint64),
        ("syncword_freq", C.c_double),
        ("syncword_amplitude", C.c_float),
        ("syncword_phase", C.c_float),
        ("syncword_freq_bin", C.c_int32),
        ("syncword_noise_power", C.c_float),
        ("syncword_esn0_db", C.c_float),
        ("syncword_time_est", C.c_float),
    ]


class StreamTag(C.Structure):
    _fields_ = [
        ("index", C.c_uint64),
        ("has_syncword", C.c_uint32),
        ("other", C.c_uint32),
        ("sw", SyncwordTag),
    ]


class Shard(C.Structure):
    _fields_ = [("device", C.c_int32), ("_pad", C.c_int32), ("first_block", C.c_uint64), ("n_blocks", C.c_uint64),
                ("total_blocks", C.c_uint64), ("first_sample", C.c_uint64), ("n_samples", C.c_uint64)]


class SfConfig(C.Structure):
    _fields_ = [
        ("samples_per_symbol", C.c_uint32),
        ("taps", C.c_void_p),
        ("n_taps", C.c_uint32),
        ("num_arms", C.c_uint32),
        ("delay", C.c_uint32),
        ("device", C.c_int32),
    ]


class ClConfig(C.Structure):
    _fields_ = [("loop_bandwidth", C.c_double), ("constellation", C.c_uint32), ("device", C.c_int32)]


class StimConfig(C.Structure):
    _fields_ = [("taps", C.c_void_p), ("n_taps", C.c_uint32), ("interpolation", C.c_uint32),
                ("syncword_symbols", C.c_void_p), ("n_syncword", C.c_uint32), ("header_symbols", C.c_uint32),
                ("payload_symbols", C.c_uint32), ("gap_symbols", C.c_uint32), ("phase_incr", C.c_float),
                ("noise_amplitude", C.c_float), ("seed", C.c_uint64), ("device", C.c_int32)]


class SdfHeader(C.Structure):
    _fields_ = [("invalid_header", C.c_uint32), ("packet_length", C.c_uint64)]


class FeConfig(C.Structure):
    _fields_ = [
        ("rate", C.c_float),
        ("taps", C.c_void_p),
        ("n_taps", C.c_uint32),
        ("filter_size", C.c_uint32),
        ("phase_incr", C.c_float),
        ("enable_resampler", C.c_uint32),
        ("enable_rotator", C.c_uint32),
        ("device", C.c_int32),
        ("rate_is_f64", C.c_uint32),
        ("rate_f64", C.c_double),
        ("fp_contract", C.c_uint32),
        ("_reserved", C.c_uint32),
    ]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C gr4_packet_modem_b200/csrc).  There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.b200sync_last_error.restype = C.c_char_p
    L.b200sync_abi_version.restype = C.c_int
    L.b200sync_launch_count.restype = C.c_uint64
    vp, sz, psz = C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)
    L.b200sync_host_register.argtypes = [vp, sz]
    L.b200sync_host_unregister.argtypes = [vp]
    L.b200sync_sd_set_auto_register.argtypes = [vp, C.c_int]
    L.b200sync_sd_create.argtypes = [C.POINTER(SdConfig), C.POINTER(vp)]
    L.b200sync_sd_destroy.argtypes = [vp]
    L.b200sync_sd_start.argtypes = [vp]
    L.b200sync_sd_info.argtypes = [vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_float),
                                   C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
    L.b200sync_sd_process.argtypes = [vp, vp, sz, vp, psz, vp, sz, psz]
    L.b200sync_sd_tags_ready.argtypes = [vp]
    L.b200sync_sd_tags_ready.restype = C.c_size_t
    L.b200sync_sd_drain_tags.argtypes = [vp, vp, sz, psz]
    L.b200sync_sd_drain_tags.restype = C.c_int
    L.b200sync_sd_detect_device.argtypes = [vp, vp, sz, vp, vp, vp, sz, psz, psz]
    L.b200sync_sd_detect_host.argtypes = [vp, vp, sz, vp, sz, psz, psz]
    L.b200sync_sd_detect_host_out.argtypes = [vp, vp, sz, vp, vp, sz, psz, psz]
    L.b200sync_sd_detect_host_out.restype = C.c_int
    L.b200sync_sd_shard_output.argtypes = [vp, vp, C.c_uint64, sz]
    L.b200sync_sd_shard_output.restype = C.c_int
    L.b200sync_sd_shard_output_host.argtypes = [vp, vp, C.c_uint64, sz]
    L.b200sync_sd_shard_output_host.restype = C.c_int
    L.b200sync_sd_detect_file.argtypes = [vp, C.c_char_p, C.c_uint64, C.c_uint64, vp, sz, psz, psz,
                                          C.POINTER(C.c_uint64)]
    L.b200sync_sd_detect_file.restype = C.c_int
    L.b200sync_sd_detect_channels_device.argtypes = [vp, vp, sz, sz, sz, vp, vp, sz, vp, psz]
    L.b200sync_sd_shard_phase1.argtypes = [vp, vp, C.c_uint64, sz, C.c_uint64, C.c_uint64, C.c_uint64, vp, vp, sz]
    L.b200sync_sd_shard_phase1_host.argtypes = [vp, vp, C.c_uint64, sz, C.c_uint64, C.c_uint64, C.c_uint64, vp, sz]
    L.b200sync_sd_shard_phase1_host.restype = C.c_int
    L.b200sync_sd_shard_phase2.argtypes = [vp, C.c_uint32, vp, sz, psz]
    L.b200sync_sd_records_to_tags.argtypes = [vp, vp, sz, vp]
    L.b200sync_sd_copy_metric.argtypes = [vp, vp, sz]
    pf = C.POINTER(C.c_float)
    L.b200sync_sd_last_timings.argtypes = [vp, pf, pf, pf]
    L.b200sync_fe_create.argtypes = [C.POINTER(FeConfig), C.POINTER(vp)]
    L.b200sync_fe_destroy.argtypes = [vp]
    L.b200sync_fe_start.argtypes = [vp]
    L.b200sync_fe_last_error.restype = C.c_char_p
    L.b200sync_fe_max_output.argtypes = [vp, sz]
    L.b200sync_fe_max_output.restype = sz
    L.b200sync_fe_process.argtypes = [vp, vp, sz, vp, sz, psz, psz]
    L.b200sync_fe_process_device.argtypes = [vp, vp, sz, vp, sz, vp, psz, psz]
    L.b200sync_sf_create.argtypes = [C.POINTER(SfConfig), C.POINTER(vp)]
    L.b200sync_sf_destroy.argtypes = [vp]
    L.b200sync_sf_start.argtypes = [vp]
    L.b200sync_sf_last_error.restype = C.c_char_p
    L.b200sync_sf_process.argtypes = [vp, vp, sz, vp, sz, vp, sz, psz, psz, vp, sz, psz]
    L.b200sync_sf_process_device.argtypes = [vp, vp, sz, vp, sz, vp, sz, vp, psz, psz, vp, sz, psz]
    L.b200sync_sf_fuse_cfc.argtypes = [vp, C.c_int, C.c_uint32]
    L.b200sync_sf_fuse_cfc.restype = C.c_int
    L.b200sync_cfc_create.argtypes = [C.c_uint32, C.c_int32, C.POINTER(vp)]
    L.b200sync_cfc_destroy.argtypes = [vp]
    L.b200sync_cfc_start.argtypes = [vp]
    L.b200sync_cfc_last_error.restype = C.c_char_p
    L.b200sync_cfc_process.argtypes = [vp, vp, sz, vp, sz, vp]
    L.b200sync_cfc_process_device.argtypes = [vp, vp, sz, vp, sz, vp, vp]
    for name in ("b200sync_cfc_create", "b200sync_cfc_start", "b200sync_cfc_process", "b200sync_cfc_process_device"):
        getattr(L, name).restype = C.c_int
    L.b200sync_wo_create.argtypes = [vp, C.c_uint32, C.c_int32, C.POINTER(vp)]
    L.b200sync_wo_destroy.argtypes = [vp]
    L.b200sync_wo_start.argtypes = [vp]
    L.b200sync_wo_process.argtypes = [vp, vp, sz, vp, sz, vp]
    L.b200sync_wo_process_device.argtypes = [vp, vp, sz, vp, sz, vp, vp]
    L.b200sync_cl_create.argtypes = [C.POINTER(ClConfig), C.POINTER(vp)]
    L.b200sync_cl_destroy.argtypes = [vp]
    L.b200sync_cl_start.argtypes = [vp]
    L.b200sync_cl_last_error.restype = C.c_char_p
    L.b200sync_cl_info.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.b200sync_cl_process.argtypes = [vp, vp, sz, vp, sz, vp]
    L.b200sync_cl_process_device.argtypes = [vp, vp, sz, vp, sz, vp, vp]
    L.b200sync_cl_fuse_wipeoff.argtypes = [vp, vp, C.c_uint32]
    L.b200sync_cl_state.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    for name in ("b200sync_wo_create", "b200sync_wo_start", "b200sync_wo_process", "b200sync_wo_process_device",
                 "b200sync_cl_create", "b200sync_cl_start", "b200sync_cl_info", "b200sync_cl_process",
                 "b200sync_cl_process_device", "b200sync_cl_fuse_wipeoff", "b200sync_cl_state"):
        getattr(L, name).restype = C.c_int
    L.b200sync_stim_create.argtypes = [C.POINTER(StimConfig), C.POINTER(vp)]
    L.b200sync_stim_create.restype = C.c_int
    L.b200sync_stim_destroy.argtypes = [vp]
    L.b200sync_stim_last_error.restype = C.c_char_p
    L.b200sync_stim_generate_device.argtypes = [vp, C.c_uint64, sz, vp, vp]
    L.b200sync_stim_generate_device.restype = C.c_int
    L.b200sync_sdf_create.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(vp)]
    L.b200sync_sdf_destroy.argtypes = [vp]
    L.b200sync_sdf_start.argtypes = [vp]
    pi = C.POINTER(C.c_int)
    u64 = C.c_uint64
    L.b200sync_sd_shard_phase1_file.argtypes = [vp, C.c_char_p, u64, u64, sz, u64, u64, u64, vp, sz]
    L.b200sync_sd_multi_create.argtypes = [C.POINTER(SdConfig), vp, sz, C.POINTER(vp)]
    L.b200sync_sd_multi_destroy.argtypes = [vp]
    L.b200sync_sd_multi_destroy.restype = None
    L.b200sync_sd_multi_devices.argtypes = [vp]
    L.b200sync_sd_multi_devices.restype = sz
    L.b200sync_sd_multi_context.argtypes = [vp, sz]
    L.b200sync_sd_multi_context.restype = vp
    L.b200sync_sd_multi_plan.argtypes = [vp, u64, C.POINTER(Shard)]
    L.b200sync_sd_multi_detect_device.argtypes = [vp, C.POINTER(vp), u64, vp, sz, psz, psz]
    L.b200sync_sd_multi_detect_host.argtypes = [vp, vp, u64, vp, sz, psz, psz]
    L.b200sync_sd_multi_detect_file.argtypes = [vp, C.c_char_p, u64, u64, vp, sz, psz, psz, C.POINTER(u64)]
    L.b200sync_sd_multi_last_timings.argtypes = [vp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    for name in ("b200sync_sd_shard_phase1_file", "b200sync_sd_multi_create", "b200sync_sd_multi_plan",
                 "b200sync_sd_multi_detect_device", "b200sync_sd_multi_detect_host", "b200sync_sd_multi_detect_file",
                 "b200sync_sd_multi_last_timings"):
        getattr(L, name).restype = C.c_int
    L.b200sync_sdf_process.argtypes = [vp, vp, sz, vp, sz, vp, vp, sz, psz, psz, psz, vp, pi, pi]
    for name in ("b200sync_sf_create", "b200sync_sf_start", "b200sync_sf_process", "b200sync_sf_process_device",
                 "b200sync_sdf_create", "b200sync_sdf_start", "b200sync_sdf_process"):
        getattr(L, name).restype = C.c_int
    for name in ("b200sync_fe_create", "b200sync_fe_start", "b200sync_fe_process", "b200sync_fe_process_device"):
        getattr(L, name).restype = C.c_int
    for name in ("b200sync_sd_create", "b200sync_sd_start", "b200sync_sd_info", "b200sync_sd_process",
                 "b200sync_sd_detect_device", "b200sync_sd_detect_host", "b200sync_sd_detect_channels_device",
                 "b200sync_sd_shard_phase1",
                 "b200sync_sd_shard_phase2", "b200sync_sd_records_to_tags", "b200sync_sd_copy_metric",
                 "b200sync_sd_last_timings"):
        getattr(L, name).restype = C.c_int
    _lib = L
    return L


class B200SyncError(RuntimeError):
    """Raised where the reference block would throw gr::exception."""


def check(rc: int) -> int:
    if rc < 0:
        raise B200SyncError(f"libb200sync error {rc}: {lib().b200sync_last_error().decode()}")
    return rc


def check_cfc(rc: int) -> int:
    if rc < 0:
        raise B200SyncError(f"libb200sync error {rc}: {lib().b200sync_cfc_last_error().decode()}")
    return rc


def check_sf(rc: int) -> int:
    if rc < 0:
        raise B200SyncError(f"libb200sync error {rc}: {lib().b200sync_sf_last_error().decode()}")
    return rc


def check_fe(rc: int) -> int:
    if rc < 0:
        raise B200SyncError(f"libb200sync error {rc}: {lib().b200sync_fe_last_error().decode()}")
    return rc


def check_cl(rc: int) -> int:
    if rc < 0:
        raise B200SyncError(f"libb200sync error {rc}: {lib().b200sync_cl_last_error().decode()}")
    return rc


def check_stim(rc: int) -> int:
    if rc < 0:
        raise B200SyncError(f"libb200sync error {rc}: {lib().b200sync_stim_last_error().decode()}")
    return rc
