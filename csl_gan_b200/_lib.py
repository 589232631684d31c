"""ctypes binding of libcslgan_b200.so -- the only door from Python into the CUDA kernels.

The binding is deliberately thin: raw `tensor.data_ptr()` integers, sizes, and the current
CUDA stream handle.  There is no CPU or eager-PyTorch fallback: if the library is missing or
a call fails, a `CslGanCudaError` is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcslgan_b200.so")
CG_MAX_KH = 16
EPI_SUMSQ, EPI_ACCUM, EPI_STORE, EPI_STORE_NATURAL = 0, 1, 2, 3
GROUP_SAMPLE, GROUP_SPLITK = 0, 1


class CslGanCudaError(RuntimeError):
    pass


class UnfoldGeom(C.Structure):
    _fields_ = [(n, C.c_int) for n in
                ("C", "H", "W", "KH", "KW", "sh", "sw", "ph", "pw", "dh", "dw", "Ho", "Wo")]


class UnfoldPlan(C.Structure):
    _fields_ = [("n_rho", C.c_int), ("Hs", C.c_int), ("Wop", C.c_int), ("rows", C.c_int), ("slot_stride", C.c_int),
                ("tap_row0", C.c_int * CG_MAX_KH), ("tap_coloff", C.c_int * CG_MAX_KH),
                ("rho", C.c_int * CG_MAX_KH), ("a_min", C.c_int)]


class GhostPlan(C.Structure):
    _fields_ = [("n_rh", C.c_int), ("n_rw", C.c_int), ("Hs", C.c_int), ("Ws", C.c_int), ("ah_min", C.c_int),
                ("aw_min", C.c_int), ("Cp", C.c_int), ("rho_h", C.c_int * CG_MAX_KH), ("rho_w", C.c_int * CG_MAX_KH),
                ("tap_plane", C.c_int * (CG_MAX_KH * CG_MAX_KH)), ("tap_hoff", C.c_int * (CG_MAX_KH * CG_MAX_KH)),
                ("tap_woff", C.c_int * (CG_MAX_KH * CG_MAX_KH)), ("slot_stride", C.c_longlong),
                ("merged", C.c_int), ("Cs", C.c_int), ("n_taps", C.c_int), ("cw", C.c_int)]


ClPlan = GhostPlan


class ClDesc(C.Structure):
    _fields_ = [("Xt", C.c_void_p), ("xt_pitch", C.c_longlong), ("xt_rows", C.c_longlong), ("M", C.c_int),
                ("Yt", C.c_void_p), ("n_slots_total", C.c_int),
                ("group_mode", C.c_int), ("n_groups", C.c_int), ("slot_lo", C.c_int), ("slot_hi", C.c_int),
                ("epi", C.c_int), ("out", C.c_void_p), ("out_group_stride", C.c_longlong), ("max_ctas", C.c_int),
                ("n_seg", C.c_int), ("seg_stride", C.c_int), ("pair", C.c_int),
                ("half", C.c_int), ("inv_x", C.c_void_p), ("inv_y", C.c_void_p), ("out_scale", C.c_void_p)]


class GhostDesc(C.Structure):
    _fields_ = [("Xt", C.c_void_p), ("xt_pitch", C.c_longlong), ("xt_rows", C.c_longlong),
                ("Yt", C.c_void_p), ("n_slots_total", C.c_int), ("O", C.c_int),
                ("slot0", C.c_int), ("n_slots", C.c_int), ("norm2", C.c_void_p), ("max_ctas", C.c_int),
                ("half", C.c_int), ("inv_x", C.c_void_p), ("inv_y", C.c_void_p)]


class ContractDesc(C.Structure):
    _fields_ = [
        ("X", C.c_void_p), ("x_pitch", C.c_longlong), ("x_rows", C.c_int), ("x_cols", C.c_longlong),
        ("Y", C.c_void_p), ("y_pitch", C.c_longlong), ("y_rows", C.c_int), ("y_cols", C.c_longlong),
        ("M", C.c_int), ("C", C.c_int), ("KH", C.c_int), ("KW", C.c_int),
        ("tap_row0", C.c_int * CG_MAX_KH), ("tap_coloff", C.c_int * CG_MAX_KH),
        ("nkb", C.c_int), ("x_slot_stride", C.c_longlong), ("y_slot_stride", C.c_longlong),
        ("group_mode", C.c_int), ("n_groups", C.c_int),
        ("slot_lo", C.c_int), ("slot_hi", C.c_int), ("spg", C.c_int),
        ("n_seg", C.c_int), ("seg_stride", C.c_int),
        ("epi", C.c_int), ("out", C.c_void_p), ("out_group_stride", C.c_longlong),
        ("block_n", C.c_int), ("max_ctas", C.c_int),
    ]


OP_ROW_SUMSQ, OP_COPY, OP_MUL, OP_WCOLSUM, OP_CLIP_MULT = 0, 1, 2, 3, 4


class SmallOp(C.Structure):
    _fields_ = [("op", C.c_int), ("R", C.c_int), ("lo", C.c_int), ("n", C.c_longlong), ("a", C.c_void_p), ("b", C.c_void_p),
                ("c", C.c_void_p), ("out", C.c_void_p), ("out2", C.c_void_p)]


class ScaleSeg(C.Structure):
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("mult", C.c_void_p), ("pitch", C.c_longlong),
                ("slot_stride", C.c_longlong), ("rows", C.c_int), ("slot_lo", C.c_int), ("slot_hi", C.c_int)]


class NoiseSeg(C.Structure):
    _fields_ = [("inp", C.c_void_p), ("grad", C.c_void_p), ("n", C.c_longlong), ("std_mult", C.c_double),
                ("std_dev", C.c_void_p)]


_PROTOS = {
    "cg_version": (C.c_int, []),
    "cg_last_error": (C.c_char_p, []),
    "cg_device_info": (C.c_int, [C.POINTER(C.c_int)] * 4),
    "cg_plan_unfold": (C.c_int, [C.POINTER(UnfoldGeom), C.POINTER(UnfoldPlan)]),
    "cg_stage_rows_t": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_longlong, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_void_p]),
    "cg_stage_rows": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p]),
    "cg_stage_unfold": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(UnfoldGeom), C.POINTER(UnfoldPlan), C.c_float,
                                  C.c_void_p, C.c_longlong, C.c_int, C.c_void_p]),
    "cg_contract": (C.c_int, [C.POINTER(ContractDesc), C.c_void_p]),
    "cg_plan_ghost": (C.c_int, [C.POINTER(UnfoldGeom), C.POINTER(GhostPlan)]),
    "cg_ghost_norm": (C.c_int, [C.POINTER(GhostDesc), C.POINTER(UnfoldGeom), C.POINTER(GhostPlan), C.c_void_p]),
    "cg_plan_cl": (C.c_int, [C.POINTER(UnfoldGeom), C.c_int, C.POINTER(GhostPlan)]),
    "cg_plan_cl_cw": (C.c_int, [C.POINTER(UnfoldGeom), C.c_int, C.c_int, C.POINTER(GhostPlan)]),
    "cg_stage_xt": (C.c_int, [C.c_void_p, C.c_longlong, C.c_longlong, C.c_longlong, C.c_longlong, C.c_int, C.c_int,
                              C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p,
                              C.c_void_p]),
    "cg_stage_yt": (C.c_int, [C.c_void_p, C.c_longlong, C.c_longlong, C.c_longlong, C.c_longlong, C.c_int,
                              C.POINTER(UnfoldGeom), C.POINTER(GhostPlan), C.c_float, C.c_void_p, C.c_int, C.c_int,
                              C.c_void_p]),
    "cg_stage_xt_h": (C.c_int, [C.c_void_p, C.c_longlong, C.c_longlong, C.c_longlong, C.c_longlong, C.c_int, C.c_int,
                                C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_void_p,
                                C.c_void_p, C.c_void_p, C.c_void_p]),
    "cg_stage_yt_h": (C.c_int, [C.c_void_p, C.c_longlong, C.c_longlong, C.c_longlong, C.c_longlong, C.c_int,
                                C.POINTER(UnfoldGeom), C.POINTER(GhostPlan), C.c_float, C.c_void_p, C.c_int, C.c_int,
                                C.c_void_p, C.c_void_p, C.c_void_p]),
    "cg_scale_slots_h_multi": (C.c_int, [C.POINTER(ScaleSeg), C.c_int, C.c_void_p]),
    "cg_small_ops": (C.c_int, [C.POINTER(SmallOp), C.c_int, C.c_void_p]),
    "cg_thin_direct_ok": (C.c_int, [C.POINTER(UnfoldGeom), C.c_int]),
    "cg_thin_capture": (C.c_int, [C.c_void_p, C.c_longlong, C.c_longlong, C.c_longlong, C.c_longlong, C.c_void_p,
                                  C.c_int, C.POINTER(UnfoldGeom), C.c_int, C.c_float, C.c_void_p, C.c_longlong,
                                  C.c_void_p, C.c_void_p, C.c_void_p]),
    "cg_thin_capture2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_longlong, C.c_longlong, C.c_void_p,
                                   C.c_void_p, C.c_int, C.c_int, C.POINTER(UnfoldGeom), C.c_int, C.c_float, C.c_void_p,
                                   C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cg_clip_mult": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cg_scale_slots_h": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_longlong, C.c_int, C.c_int,
                                   C.c_void_p, C.c_void_p]),
    "cg_outer_rows_cl_h": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "cg_cl_contract": (C.c_int, [C.POINTER(ClDesc), C.POINTER(UnfoldGeom), C.POINTER(GhostPlan), C.c_void_p]),
    "cg_cl_pair_ok": (C.c_int, [C.c_int, C.POINTER(UnfoldGeom), C.POINTER(GhostPlan)]),
    "cg_cl_kblock_rows": (C.c_int, [C.POINTER(UnfoldGeom), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "cg_rowpair_dot": (C.c_int, [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                 C.c_void_p]),
    "cg_joint_rows_sumsq": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]),
    "cg_outer_rows_cl": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_void_p, C.c_void_p]),
    "cg_outer_rows": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_void_p, C.c_void_p]),
    "cg_row_sumsq": (C.c_int, [C.c_void_p, C.c_longlong, C.c_longlong, C.c_longlong, C.c_void_p, C.c_int, C.c_void_p]),
    "cg_vec_mul": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p]),
    "cg_vec_fma": (C.c_int, [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_longlong, C.c_void_p]),
    "cg_clip_factors": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_int, C.c_int,
                                  C.c_void_p, C.c_void_p, C.c_void_p]),
    "cg_scale_slots": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_longlong, C.c_int, C.c_int,
                                 C.c_void_p, C.c_void_p]),
    "cg_permute_accum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "cg_weighted_colsum": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "cg_row_stat": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    "cg_noise_finalize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_double, C.c_double, C.c_double,
                                    C.c_ulonglong, C.c_ulonglong, C.POINTER(C.c_ulonglong), C.c_void_p]),
    "cg_noise_finalize_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_double, C.c_double, C.c_void_p,
                                        C.c_double, C.c_ulonglong, C.c_ulonglong, C.POINTER(C.c_ulonglong), C.c_void_p]),
    "cg_noise_finalize_graph": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_double, C.c_double, C.c_void_p,
                                          C.c_double, C.c_ulonglong, C.c_void_p, C.c_ulonglong,
                                          C.POINTER(C.c_ulonglong), C.c_void_p]),
    "cg_philox_advance": (C.c_int, [C.c_void_p, C.c_ulonglong, C.c_void_p]),
    "cg_noise_finalize_multi": (C.c_int, [C.POINTER(NoiseSeg), C.c_int, C.c_double, C.c_void_p, C.c_double, C.c_void_p,
                                          C.c_ulonglong, C.c_ulonglong, C.c_void_p, C.POINTER(C.c_ulonglong),
                                          C.c_void_p]),
    "cg_noise_finalize_allreduce": (C.c_int, [C.POINTER(NoiseSeg), C.c_int, C.c_int, C.c_ulonglong, C.c_ulonglong, C.c_void_p,
                                              C.POINTER(C.c_ulonglong), C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong,
                                              C.c_longlong, C.c_int, C.c_int, C.c_void_p]),
    "cg_row_l2_norm": (C.c_int, [C.c_void_p, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p]),
    "cg_row_l2_norm_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p]),
    "cg_vec_max": (C.c_int, [C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p]),
    "cg_l2_clip": (C.c_int, [C.c_void_p, C.c_longlong, C.c_longlong, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_PROTOS)

_lib: Optional[C.CDLL] = None
launch_count = 0          # number of ABI calls that enqueue GPU work (bench.py reports it)
_NO_LAUNCH = {"cg_version", "cg_last_error", "cg_device_info", "cg_plan_unfold", "cg_plan_ghost", "cg_plan_cl",
              "cg_plan_cl_cw", "cg_cl_pair_ok", "cg_cl_kblock_rows", "cg_thin_direct_ok"}


def load() -> C.CDLL:
    """Load the shared library (building is `python -m csl_gan_b200.build`, done by
    __graft_entry__.build()).  Fails loudly: there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise CslGanCudaError(
            f"{LIB_PATH} is missing: build it with `python -m csl_gan_b200.build` "
            "(the DP hot path has no CPU / eager fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


_profile = None           # when a list: (name, start_event, end_event) per launching ABI call


def set_profile(on: bool):
    """Bracket every launching ABI call with CUDA events on the current stream (bench.py uses this
    to attribute device time to kernels without a profiler attached)."""
    global _profile
    _profile = [] if on else None


def profile_summary():
    """{abi call: (total ms, launches)}; call after torch.cuda.synchronize()."""
    out = {}
    for name, e0, e1 in _profile or []:
        ms, n = out.get(name, (0.0, 0))
        out[name] = (ms + e0.elapsed_time(e1), n + 1)
    return out


def call(name: str, *args):
    global launch_count
    lib = load()
    launching = name not in _NO_LAUNCH
    if _profile is not None and launching:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args)
        e1.record()
        _profile.append((name, e0, e1))
    else:
        rc = getattr(lib, name)(*args)
    if rc != 0:
        raise CslGanCudaError(f"{name}: {lib.cg_last_error().decode(errors='replace')}")
    if launching:
        launch_count += 1


def noise_multi(segs, in_div, in_div_dev, noise_div, noise_div_dev, seed, offset, offset_dev, stream) -> int:
    """One launch of cg_noise_finalize_multi over `segs` = [(in tensor or None, grad tensor, std_mult, std_dev
    tensor or None)]; returns the generator-offset advance."""
    arr = (NoiseSeg * len(segs))()
    for i, (tin, tg, mult, sdev) in enumerate(segs):
        arr[i].inp = ptr(tin)
        arr[i].grad = ptr(tg)
        arr[i].n = tg.numel()
        arr[i].std_mult = float(mult)
        arr[i].std_dev = ptr(sdev)
    inc = C.c_ulonglong(0)
    call("cg_noise_finalize_multi", arr, len(segs), float(in_div), ptr(in_div_dev), float(noise_div), ptr(noise_div_dev),
         int(seed), int(offset), ptr(offset_dev), C.byref(inc), stream)
    return inc.value


def noise_multi_allreduce(segs, mean: bool, seed, offset, offset_dev, local_flat: torch.Tensor, mc_ptr: int,
                          peer_ptrs, count_off: int, rank: int, world: int, stream) -> int:
    """One launch of cg_noise_finalize_allreduce: cross-rank sum + noise + broadcast over the multicast mapping
    `mc_ptr` of the symmetric buffer `local_flat` (or, mc_ptr = 0, over the peer mappings `peer_ptrs`); `segs` as in
    noise_multi.  Returns the generator-offset advance."""
    peers = None
    if peer_ptrs:
        peers = (C.c_void_p * len(peer_ptrs))(*[int(x) for x in peer_ptrs])
    arr = (NoiseSeg * len(segs))()
    for i, (tin, tg, mult, sdev) in enumerate(segs):
        arr[i].inp = ptr(tin)
        arr[i].grad = ptr(tg)
        arr[i].n = tg.numel()
        arr[i].std_mult = float(mult)
        arr[i].std_dev = ptr(sdev)
    inc = C.c_ulonglong(0)
    call("cg_noise_finalize_allreduce", arr, len(segs), 1 if mean else 0, int(seed), int(offset), ptr(offset_dev),
         C.byref(inc), local_flat.data_ptr(), int(mc_ptr) or None, peers, local_flat.numel(), int(count_off), int(rank),
         int(world), stream)
    return inc.value


def small_op(op: int, a, out, n: int, b=None, c=None, out2=None, R: int = 0, lo: int = 0):
    """One entry of a cg_small_ops table; tensors (or None), kept alive by the caller until the launch."""
    return (op, R, lo, n, a, b, c, out, out2)


def small_ops(ops, stream):
    """Launch a table of small operations (see include/cslgan_b200.h cg_small_ops) in batches of 32."""
    for i in range(0, len(ops), 32):
        chunk = ops[i:i + 32]
        arr = (SmallOp * len(chunk))()
        for j, (op, R, lo, n, a, b, c, out, out2) in enumerate(chunk):
            arr[j].op, arr[j].R, arr[j].lo, arr[j].n = op, int(R), int(lo), int(n)
            arr[j].a, arr[j].b, arr[j].c, arr[j].out, arr[j].out2 = ptr(a), ptr(b), ptr(c), ptr(out), ptr(out2)
        call("cg_small_ops", arr, len(chunk), stream)


def scale_slots_multi(segs, stream):
    """One launch of cg_scale_slots_h_multi over segs = [(src, dst, mult, pitch, slot_stride, rows, slot_lo, slot_hi)]."""
    for i in range(0, len(segs), 8):
        chunk = segs[i:i + 8]
        arr = (ScaleSeg * len(chunk))()
        for j, (src, dst, mult, pitch, stride, rows, lo, hi) in enumerate(chunk):
            arr[j].src, arr[j].dst, arr[j].mult = ptr(src), ptr(dst), ptr(mult)
            arr[j].pitch, arr[j].slot_stride, arr[j].rows, arr[j].slot_lo, arr[j].slot_hi = int(pitch), int(stride), int(rows), int(lo), int(hi)
        call("cg_scale_slots_h_multi", arr, len(chunk), stream)


def cl_pair_ok(M: int, geom, plan) -> bool:
    """May cg_cl_contract run this layer's clipped sum on CTA pairs (cg_cl_desc.pair = 1)?"""
    return bool(load().cg_cl_pair_ok(int(M), C.byref(geom), C.byref(plan)))


def thin_direct_ok(geom, M: int) -> bool:
    """Can cg_thin_capture compute this layer's per-sample gradients straight from the critic's tensors?"""
    return bool(load().cg_thin_direct_ok(C.byref(geom), int(M)))


def stream_ptr(device=None) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    return t.data_ptr()


def require_cuda_f32(t: torch.Tensor, what: str) -> torch.Tensor:
    if not t.is_cuda:
        raise CslGanCudaError(f"{what} must live on a CUDA device (got {t.device}); there is no CPU path")
    if t.dtype != torch.float32:
        raise CslGanCudaError(f"{what} must be float32 (got {t.dtype})")
    return t.contiguous()


def plan_unfold(C_, H, W, KH, KW, sh, sw, ph, pw, dh, dw, Ho, Wo):
    g = UnfoldGeom(C_, H, W, KH, KW, sh, sw, ph, pw, dh, dw, Ho, Wo)
    p = UnfoldPlan()
    call("cg_plan_unfold", C.byref(g), C.byref(p))
    return g, p


def plan_ghost(geom: UnfoldGeom):
    """GhostPlan for the geometry, or None when it is outside the ghost-norm envelope."""
    lib = load()
    p = GhostPlan()
    if lib.cg_plan_ghost(C.byref(geom), C.byref(p)) != 0:
        return None
    return p


def plan_cl(geom: UnfoldGeom, merged: bool, cw: int = 32) -> GhostPlan:
    """Staging plan of the channels-last path; cw = channels per 128-byte chunk row (32 TF32 words / 64 FP16)."""
    p = GhostPlan()
    call("cg_plan_cl_cw", C.byref(geom), int(merged), cw, C.byref(p))      # merged: 0 / 1 (kw) / 2 (kh and kw)
    return p


def default_operand_dtype() -> str:
    """Operand containers of the channels-last contraction path: "f16" (default: FP16 with an exact per-sample
    power-of-two scale; TF32's mantissa, half the bytes, twice the tensor-core rate) or "tf32"
    (environment CSLGAN_OPERANDS overrides, for A/B measurements)."""
    v = os.environ.get("CSLGAN_OPERANDS", "f16").lower()
    if v not in ("f16", "tf32"):
        raise CslGanCudaError(f"CSLGAN_OPERANDS must be f16 or tf32, got {v!r}")
    return v


def cl_kblock_rows(geom, half: bool):
    """(contraction rows per k-block, slots per k-block) of the split-K clipped sum for this window grid."""
    r, ks = C.c_int(), C.c_int()
    call("cg_cl_kblock_rows", C.byref(geom), 1 if half else 0, C.byref(r), C.byref(ks))
    return r.value, ks.value


def cl_supported(Ho: int, Wo: int) -> bool:
    """Mirror of cl_kblock() in csrc/abi.cu: can the window grid be tiled into 32-position k-blocks?"""
    Q = Ho * Wo
    if Q >= 32:
        if Q % 32:
            return False
        w = min(Wo, 32)
        return Wo % w == 0 and 32 % w == 0 and Ho % (32 // w) == 0
    return 32 % Q == 0


def device_info():
    vals = [C.c_int() for _ in range(4)]
    call("cg_device_info", *[C.byref(v) for v in vals])
    return tuple(v.value for v in vals)
