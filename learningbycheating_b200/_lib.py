"""ctypes binding of include/lbc_b200.h (the C ABI of the native hot path).

The package loads exactly one library: ``liblbc_b200.so`` next to this file (built in-tree by
``learningbycheating_b200/build.py`` with nvcc for sm_100a).  If it is missing or no CUDA device is
usable the package raises -- there is no CPU / PyTorch fallback.  ``use_library_for_tests`` exists
only so the ``-m "not gpu"`` unit tests can point the binding at the host-emulation build of the same
sources (tests/hostemu/), which exercises the host-side logic on machines without a GPU.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblbc_b200.so")
_lib = None
_host_emu = False

c_f32p = ctypes.c_void_p
c_i64 = ctypes.c_int64


class LbcError(RuntimeError):
    pass


def _declare(lib):
    vp, i, f, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_int64
    sigs = {
        "lbc_last_error": (ctypes.c_char_p, []),
        "lbc_device_kind": (i, []),
        "lbc_build_info": (ctypes.c_char_p, []),
        "lbc_set_fast_kernels": (i, [i]),
        "lbc_set_schedule": (i, [i, i]),
        "lbc_op_copy_channels": (i, [vp, vp, i64, i, i, vp, i, i, vp]),
        "lbc_stem_layout": (i, [i, i, i]),
        "lbc_kernel_launch_count": (ctypes.c_longlong, []),
        "lbc_prof_enable": (i, [i]),
        "lbc_prof_reset": (i, []),
        "lbc_prof_get": (i, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_longlong),
                             ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]),
        "lbc_net_create": (i, [i, i, i, ctypes.POINTER(vp)]),
        "lbc_net_destroy": (None, [vp]),
        "lbc_net_num_params": (i, [vp]),
        "lbc_net_param_info": (i, [vp, i, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(i), ctypes.POINTER(i * 4),
                                   ctypes.POINTER(i64), ctypes.POINTER(i64), ctypes.POINTER(i)]),
        "lbc_net_num_buffers": (i, [vp]),
        "lbc_net_buffer_info": (i, [vp, i, ctypes.POINTER(ctypes.c_char_p), ctypes.POINTER(i64), ctypes.POINTER(i64)]),
        "lbc_net_total_params": (i64, [vp]),
        "lbc_net_total_buffers": (i64, [vp]),
        "lbc_net_workspace_bytes": (i64, [vp]),
        "lbc_net_bind": (i, [vp, vp, vp, vp]),
        "lbc_net_forward": (i, [vp, vp, vp, vp, i, i, vp, vp, vp]),
        "lbc_net_forward_u8": (i, [vp, vp, i, vp, vp, i, i, vp, vp, vp]),
        "lbc_net_infer": (i, [vp, vp, vp, i, vp, vp, i, i, vp, vp, vp]),
        "lbc_net_infer_replays": (i, [vp]),
        "lbc_net_backward": (i, [vp, vp, vp, vp]),
        "lbc_net_read_tap": (i64, [vp, ctypes.c_char_p, vp, i64, vp]),
        "lbc_net_num_grad_buckets": (i, [vp]),
        "lbc_net_grad_bucket": (i, [vp, i, ctypes.POINTER(i64), ctypes.POINTER(i64)]),
        "lbc_net_enable_grad_events": (i, [vp, i]),
        "lbc_net_stream_wait_grads": (i, [vp, i, vp]),
        "lbc_phase0_target": (i, [vp, vp, i64, f, f, f, f, f, vp]),
        "lbc_l1_loss": (i, [vp, vp, i, i, f, f, f, f, f, vp, vp, vp, vp]),
        "lbc_phase1_convert_fwd": (i, [vp, vp, i64, f, f, f, f, f, vp]),
        "lbc_phase1_convert_bwd": (i, [vp, vp, vp, i64, f, f, f, f, f, vp]),
        "lbc_phase2_weight": (i, [vp, vp, vp, i, vp]),
        "lbc_adam_step": (i, [vp, vp, vp, vp, i64, f, f, f, f, i, f, vp]),
        "lbc_trace_enable": (i, [i]),
        "lbc_trace_dump": (i, [ctypes.c_char_p, i]),
        "lbc_op_conv_fwd": (i, [vp, vp, vp, i, i, i, i, i, i, i, i, i, vp, vp, vp]),
        "lbc_op_conv_dgrad": (i, [vp, vp, vp, i, i, i, i, i, i, i, i, i, vp, i, vp]),
        "lbc_op_block_dgrad_ds": (i, [vp, vp, vp, vp, vp, i, i, i, i, i, i, vp]),
        "lbc_op_conv_wgrad": (i, [vp, vp, vp, i, i, i, i, i, i, i, i, i, vp]),
        "lbc_op_bn_train": (i, [vp, vp, vp, vp, i, vp, vp, vp, i64, i, i, vp, vp, vp, vp, vp]),
        "lbc_op_bn_bwd": (i, [vp, vp, vp, vp, vp, vp, i64, i, i, vp, vp, i, i, vp]),
        "lbc_op_ew": (i, [vp, vp, vp, i64, i, i, i, vp]),
        "lbc_op_resid_bn_bwd": (i, [vp, vp, vp, vp, vp, vp, vp, vp, vp, i64, i, i, vp]),
        "lbc_op_maxpool": (i, [vp, vp, vp, vp, i, i, i, i, vp]),
        "lbc_op_bn_relu_maxpool": (i, [vp, vp, vp, vp, vp, vp, vp, vp, i, i, i, i, i, vp]),
        "lbc_op_stem_tail": (i, [vp, vp, vp, vp, vp, vp, vp, vp, i, i, i, i, i, vp]),
        "lbc_op_spatial_softmax": (i, [vp, vp, i, i, i, i, vp]),
        "lbc_op_head": (i, [vp, vp, vp, vp, vp, i, i, i, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp,
                            ctypes.POINTER(i), i, vp]),
        "lbc_op_stem": (i, [vp, vp, i, vp, i, i, i, i, i, vp, vp, vp, vp, vp, i, vp]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)   # AttributeError here == the library does not export the header's symbol
        fn.restype = res
        fn.argtypes = args
    return sorted(sigs)


EXPORTED_SYMBOLS = None


def lib():
    """The loaded native library; raises LbcError when the CUDA build is absent."""
    global _lib, EXPORTED_SYMBOLS
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise LbcError(
                "native library %s not found -- run `python -m learningbycheating_b200.build cuda` "
                "(or __graft_entry__.build()); there is no CPU fallback" % _LIB_PATH)
        l = ctypes.CDLL(_LIB_PATH)
        EXPORTED_SYMBOLS = _declare(l)
        _lib = l
    return _lib


def use_library_for_tests(path):
    """TESTS ONLY: bind to the host-emulation build (device kind 0) instead of the CUDA library."""
    global _lib, _host_emu, EXPORTED_SYMBOLS
    l = ctypes.CDLL(path)
    EXPORTED_SYMBOLS = _declare(l)
    if l.lbc_device_kind() != 0:
        raise LbcError("use_library_for_tests expects the host-emulation build")
    _lib = l
    _host_emu = True


def is_host_emulation():
    return _host_emu


def check(rc):
    if rc != 0:
        raise LbcError(lib().lbc_last_error().decode("utf-8", "replace"))


def ptr(t):
    """Device pointer of a tensor (None -> NULL); validates device / dtype / contiguity."""
    if t is None:
        return None
    import torch
    if not t.is_contiguous():
        raise LbcError("tensor must be contiguous")
    if _host_emu:
        if t.is_cuda:
            raise LbcError("host-emulation test build takes CPU tensors")
    elif not t.is_cuda:
        raise LbcError("liblbc_b200 takes CUDA tensors only (got a %s tensor); no CPU path exists" % t.device)
    return ctypes.c_void_p(t.data_ptr())


def set_schedule(wgrad_overlap=-1, pdl=-1):
    """Launch schedule of the bf16 mode (include/lbc_b200.h: lbc_set_schedule); negative = keep."""
    check(lib().lbc_set_schedule(int(wgrad_overlap), int(pdl)))


def trace(on):
    """Start (clearing) / stop recording which kernel families the library launches."""
    lib().lbc_trace_enable(1 if on else 0)


def trace_counts():
    """{kernel family: launches} since trace(True)."""
    n = lib().lbc_trace_dump(None, 0)
    buf = ctypes.create_string_buffer(n + 1)
    lib().lbc_trace_dump(buf, n + 1)
    out = {}
    for line in buf.value.decode().splitlines():
        name, _, cnt = line.rpartition("\t")
        out[name] = int(cnt)
    return out


def stream_ptr(device=None):
    if _host_emu:
        return None
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)
