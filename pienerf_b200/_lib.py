"""ctypes binding of the C-ABI in include/pienerf_b200.h.

There is deliberately NO fallback: if pienerf_b200/lib/libpienerf_b200.so is missing the import
fails loudly (build it with `make` or `__graft_entry__.build()`), and every op raises on a
non-zero return code with the library's own message.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# PN_LIB: developer override used to A/B kernel variants built with `make variant TAG=... EXTRA=-D...` (same C-ABI)
LIB_PATH = os.environ.get("PN_LIB") or os.path.join(_HERE, "lib", "libpienerf_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"pienerf_b200: CUDA library not built ({LIB_PATH} missing). Run `make` at the repo root "
        "(nvcc -gencode arch=compute_100a,code=sm_100a); there is no CPU fallback.")

lib = C.CDLL(LIB_PATH)

vp, u32, i32, f32, f64, u64 = C.c_void_p, C.c_uint32, C.c_int, C.c_float, C.c_double, C.c_uint64


class FieldT(C.Structure):
    _fields_ = [("embeddings", vp), ("offsets", vp), ("S", f32), ("H", u32), ("L", u32), ("bound", f32),
                ("w_sigma0", vp), ("w_sigma1", vp), ("w_color0", vp), ("w_color1", vp), ("w_color2", vp)]


class DeformT(C.Structure):
    _fields_ = [("p_def", vp), ("p_ori", vp), ("F_IP", vp), ("dF_IP", vp), ("n_vtx", i32), ("IP_dx", f32),
                ("density_bitfield", vp), ("bound", f32), ("cascade", u32), ("grid_size", u32),
                ("min_near", f32), ("density_scale", f32), ("dt_gamma", f32), ("max_steps", u32), ("T_thresh", f32),
                ("max_iter_num", i32), ("hgs", f32), ("cut", i32), ("cut_bounds", f32 * 6), ("num_seek_IP", i32),
                ("bg_color", f32)]


class FrameIoT(C.Structure):
    _fields_ = [("pix", vp), ("epoch", vp), ("wait_flag", vp), ("n_wait", i32), ("signal_flag", vp), ("n_signal", i32),
                ("status", vp), ("timeout_ms", u32), ("noises", vp), ("max_passes", i32), ("flags", u32)]


class QgmlsStepT(C.Structure):
    _fields_ = [("n_ip", i32), ("n_k", i32), ("iters", i32), ("dt", f64), ("dx", f64),
                ("topo", vp), ("mu", vp), ("lam", vp), ("dNx", vp), ("adj_bgn", vp), ("adj", vp), ("adj_slices", i32),
                ("Ainv", vp), ("M", vp), ("A", vp), ("active", vp), ("pcg_iters", i32),
                ("dof_rest", vp), ("dof_f", vp), ("rhs_rest", vp), ("rhs_gravity", vp),
                ("dof", vp), ("dof_vel", vp), ("scratch", vp)]


_PROTOS = {
    "pn_last_error": (C.c_char_p, []),
    "pn_version": (i32, []),
    "pn_device_sm_count": (i32, [vp]),
    "pn_grid_encode_forward": (i32, [vp, vp, vp, vp, u32, u32, u32, u32, f32, u32, vp, u32, i32, u32, i32, vp]),
    "pn_grid_encode_backward": (i32, [vp, vp, vp, vp, vp, u32, u32, u32, u32, f32, u32, vp, vp, u32, i32, u32, i32, vp]),
    "pn_grad_total_variation": (i32, [vp, vp, vp, vp, f32, u32, u32, u32, u32, f32, u32, u32, i32, i32, vp]),
    "pn_sh_encode_forward": (i32, [vp, vp, u32, u32, u32, vp, vp]),
    "pn_sh_encode_backward": (i32, [vp, vp, u32, u32, u32, vp, vp, vp]),
    "pn_near_far_from_aabb": (i32, [vp, vp, vp, u32, f32, vp, vp, vp]),
    "pn_sph_from_ray": (i32, [vp, vp, f32, u32, vp, vp]),
    "pn_morton3D": (i32, [vp, u32, vp, vp]),
    "pn_morton3D_invert": (i32, [vp, u32, vp, vp]),
    "pn_packbits": (i32, [vp, u32, f32, vp, vp]),
    "pn_march_rays": (i32, [u32, u32, vp, vp, vp, vp, f32, f32, u32, u32, u32, vp, vp, vp, vp, vp, vp, vp, vp]),
    "pn_composite_rays": (i32, [u32, u32, f32, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "pn_march_rays_quadratic_bending": (i32, [vp, vp, vp, i32, i32, vp, vp, vp, vp, i32, vp, vp, f32, vp, i32, f32, i32, vp,
                                              u32, u32, vp, vp, vp, vp, f32, f32, u32, u32, u32, vp, vp, vp, vp, vp, vp, vp, vp]),
    "pn_march_rays_train": (i32, [vp, vp, vp, f32, f32, u32, u32, u32, u32, u32, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    "pn_set_train_block_skip": (i32, [i32]),
    "pn_set_train_write_mode": (i32, [i32]),
    "pn_composite_rays_train_forward": (i32, [vp, vp, vp, vp, u32, u32, f32, vp, vp, vp, vp]),
    "pn_composite_rays_train_backward": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, u32, u32, f32, vp, vp, vp]),
    "pn_get_rays": (i32, [vp, f32, f32, f32, f32, u32, u32, vp, vp, vp]),
    "pn_build_ip_grid": (i32, [vp, i32, vp, f32, vp, i32, vp, vp, vp, vp]),
    "pn_ip_bbox": (i32, [vp, i32, f32, i32, f32, vp, vp, vp, vp]),
    "pn_field_forward": (i32, [C.POINTER(FieldT), vp, vp, u32, vp, vp, i32, vp]),
    "pn_mlp_workspace_bytes": (u64, [u32]),
    "pn_mlp_forward": (i32, [C.POINTER(FieldT), vp, vp, u32, vp, vp, vp, u64, vp]),
    "pn_render_deformed": (i32, [C.POINTER(FieldT), C.POINTER(DeformT), vp, vp, u32, vp, vp, vp, vp, vp, u64, vp, i32, vp]),
    "pn_render_deformed_ex": (i32, [C.POINTER(FieldT), C.POINTER(DeformT), vp, vp, u32, vp, vp, vp, vp, vp, u64, vp, i32, C.POINTER(FrameIoT), vp]),
    "pn_get_rays_pix": (i32, [vp, u32, u32, vp, u32, vp, vp, vp]),
    "pn_peer_alloc": (i32, [u64, C.POINTER(vp)]),
    "pn_peer_free": (i32, [vp]),
    "pn_peer_export": (i32, [vp, C.c_char_p]),
    "pn_peer_open": (i32, [C.c_char_p, C.POINTER(vp)]),
    "pn_peer_close": (i32, [vp]),
    "pn_peer_put": (i32, [vp, u64, vp, i32, vp]),
    "pn_epoch_wait": (i32, [vp, i32, vp, i32, u32, vp, u32, vp]),
    "pn_epoch_signal": (i32, [vp, vp, i32, vp]),
    "pn_render_workspace_bytes": (u64, [u32, i32, f32, f32]),
    "pn_set_render_sm_reserve": (i32, [i32]),
    "pn_set_wave_capacity": (i32, [i32]),
    "pn_set_prep_mode": (i32, [i32]),
    "pn_set_profile_events": (i32, [vp, vp]),
    "pn_set_profile_event_list": (i32, [vp, i32]),
    "pn_render_pass_count": (i32, [u32]),
    "pn_qgmls_shape_functions": (i32, [f64, vp, vp, vp, i32, vp, vp, vp, vp, vp]),
    "pn_qgmls_collect_param": (i32, [vp, vp, vp, vp, i32, i32, f64, vp, vp, vp, vp]),
    "pn_qgmls_build_ip_global": (i32, [f64, f64, vp, vp, vp, vp, vp, vp, vp, i32, i32, vp, vp]),
    "pn_qgmls_build_pin_global": (i32, [f64, vp, i32, vp, vp, i32, vp, vp]),
    "pn_qgmls_collect_gravity": (i32, [f64, vp, vp, vp, vp, i32, vp, vp]),
    "pn_qgmls_build_rhs": (i32, [f64, vp, vp, vp, vp, vp, i32, i32, vp, vp, i32, vp, vp, vp, vp]),
    "pn_qgmls_matvec3": (i32, [vp, vp, i32, vp, vp]),
    "pn_qgmls_step_scratch_doubles": (u64, [i32, i32, i32]),
    "pn_qgmls_step": (i32, [C.POINTER(QgmlsStepT), i32, vp]),
    "pn_qgmls_step_mode": (i32, [i32]),
    "pn_qgmls_step_launches": (i32, [i32, i32, i32, i32, i32]),
    "pn_qgmls_ip_info": (i32, [vp, vp, vp, vp, vp, i32, vp, vp, vp, vp]),
    "pn_qgmls_update_pos": (i32, [vp, vp, vp, i32, vp, vp]),
    "pn_qgmls_update_force": (i32, [i32, vp, vp, vp, vp, f64, i32, vp, vp]),
}

EXPORTS = tuple(_PROTOS)

for _name, (_res, _args) in _PROTOS.items():
    _fn = getattr(lib, _name)          # AttributeError here = header and library disagree: fail loudly
    _fn.restype = _res
    _fn.argtypes = _args

PN_ENOTIMPL = -3
PN_IO_WEIGHTS_READY = 1


def last_error():
    return (lib.pn_last_error() or b"").decode()


def check(rc):
    """Turn a C-ABI return code into the exceptions the reference's bindings raise."""
    if rc == 0:
        return
    msg = last_error()
    if rc == PN_ENOTIMPL:
        raise NotImplementedError(msg)
    raise RuntimeError(msg or f"pienerf_b200 error {rc}")


# ---- tensor plumbing (torch is only the device-memory / stream provider) --------------------------

def stream_ptr():
    import torch
    return vp(torch.cuda.current_stream().cuda_stream)


def _chk(t, name, dtype=None, cuda=True, contiguous=True):
    import torch
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if cuda and not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if contiguous and not t.is_contiguous():
        raise RuntimeError(f"{name} must be a contiguous tensor")
    if dtype is not None and t.dtype not in (dtype if isinstance(dtype, tuple) else (dtype,)):
        kind = "an int" if dtype == torch.int32 else "a floating" if dtype in (torch.float32, torch.float16) else str(dtype)
        raise RuntimeError(f"{name} must be {kind} tensor")


def dptr(t, name="tensor", dtype=None):
    """Device pointer of a checked tensor (None -> NULL; an int is taken as a raw device address, e.g. peer memory)."""
    if t is None:
        return vp(0)
    if isinstance(t, int):
        return vp(t)
    _chk(t, name, dtype)
    return vp(t.data_ptr())
