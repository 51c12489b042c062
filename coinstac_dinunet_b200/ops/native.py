"""ctypes binding of ``_b200_ops.so`` (C ABI ``coinn_*``; see ``csrc/*.cu``)."""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, '_b200_ops.so')

MAX_RANKS = 8


class FusedArgs(C.Structure):
    """Mirror of ``coinn::FusedArgs`` (csrc/fused_reduce_opt.cu)."""
    _fields_ = [
        ('grad_ptrs', C.c_void_p * MAX_RANKS), ('param_ptrs', C.c_void_p * MAX_RANKS),
        ('shadow_ptrs', C.c_void_p * MAX_RANKS), ('flag_ptrs', C.c_void_p * MAX_RANKS),
        ('grad_mc', C.c_void_p), ('param_mc', C.c_void_p), ('shadow_mc', C.c_void_p),
        ('m', C.c_void_p), ('v', C.c_void_p), ('epoch', C.c_void_p), ('step', C.c_void_p),
        ('ticket', C.c_void_p), ('lr_ptr', C.c_void_p), ('grad32', C.c_void_p), ('error', C.c_void_p),
        ('offset', C.c_longlong), ('numel', C.c_longlong),
        ('lr', C.c_float), ('beta1', C.c_float), ('beta2', C.c_float), ('eps', C.c_float),
        ('weight_decay', C.c_float), ('grad_scale', C.c_float), ('momentum', C.c_float),
        ('rank', C.c_int), ('world', C.c_int), ('variant', C.c_int), ('opt_kind', C.c_int),
        ('grad_dtype', C.c_int), ('zero_grads', C.c_int), ('bump_step', C.c_int), ('nesterov', C.c_int),
        ('timeout_ms', C.c_uint), ('_pad', C.c_int),
    ]


class AllGatherArgs(C.Structure):
    """Mirror of ``coinn::AllGatherArgs`` (csrc/lowrank.cu)."""
    _fields_ = [('src_ptrs', C.c_void_p * MAX_RANKS), ('flag_ptrs', C.c_void_p * MAX_RANKS), ('dst', C.c_void_p),
                ('epoch', C.c_void_p), ('error', C.c_void_p), ('numel', C.c_longlong), ('rank', C.c_int), ('world', C.c_int),
                ('timeout_ms', C.c_uint), ('_pad', C.c_int)]


class _Lib:
    def __init__(self):
        self.dll = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        d = self.dll
        d.coinn_fused_reduce_opt.argtypes = [C.POINTER(FusedArgs), C.c_int, C.c_void_p]
        d.coinn_fused_args_size.restype = C.c_int
        if d.coinn_fused_args_size() != C.sizeof(FusedArgs):
            raise RuntimeError(f'FusedArgs ABI mismatch: C {d.coinn_fused_args_size()} vs py {C.sizeof(FusedArgs)}')
        d.coinn_softmax_nll_fwd.argtypes = [C.c_void_p] * 5 + [C.c_int] * 3 + [C.c_void_p]
        d.coinn_softmax_nll_bwd.argtypes = [C.c_void_p] * 4 + [C.c_int] * 3 + [C.c_void_p]
        d.coinn_count.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_longlong,
                                  C.c_int, C.c_int, C.c_void_p]
        d.coinn_orthogonalize.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p]
        d.coinn_gemm_bf16_tn.argtypes = [C.c_void_p] * 4 + [C.c_int] * 10 + [C.c_void_p]
        d.coinn_bn_stats.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p]
        d.coinn_bn_finalize.argtypes = [C.c_void_p] * 5 + [C.c_float] * 3 + [C.c_int, C.c_void_p]
        d.coinn_bn_finalize2.argtypes = [C.c_void_p] * 5 + [C.c_float] * 3 + [C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        d.coinn_conv_block_grad_finalize.argtypes = [C.c_void_p] * 5 + [C.c_int] * 3 + [C.c_void_p]
        d.coinn_bn_relu_pool_fwd.argtypes = [C.c_void_p] * 6 + [C.c_int] * 5 + [C.c_void_p]
        d.coinn_bn_relu_pool_bwd.argtypes = [C.c_void_p] * 8 + [C.c_int] * 6 + [C.c_void_p]
        d.coinn_conv3d_igemm.argtypes = [C.c_void_p] * 3 + [C.c_int] * 7 + [C.c_void_p]
        d.coinn_conv3d_tma.argtypes = [C.c_void_p] * 3 + [C.c_int] * 7 + [C.c_void_p]
        d.coinn_conv1_padded_shape.argtypes = [C.c_int] * 4 + [C.POINTER(C.c_longlong), C.POINTER(C.c_int)]
        d.coinn_conv1_pad_input_hd.argtypes = [C.c_void_p, C.c_int, C.c_void_p] + [C.c_int] * 4 + [C.c_void_p]
        d.coinn_conv1_fused_stats.argtypes = [C.c_void_p] * 3 + [C.c_int] * 4 + [C.c_void_p]
        d.coinn_conv1_fused_pool.argtypes = [C.c_void_p] * 8 + [C.c_int] * 4 + [C.c_void_p]
        d.coinn_conv1_fused_bwd.argtypes = [C.c_void_p] * 9 + [C.c_int] * 4 + [C.c_void_p]
        d.coinn_conv3d_halo.argtypes = [C.c_void_p] * 3 + [C.c_int] * 8 + [C.c_void_p]
        d.coinn_conv3d_halo_stats.argtypes = [C.c_void_p] * 4 + [C.c_int] * 8 + [C.c_void_p]
        d.coinn_conv3d_wgrad_halo.argtypes = [C.c_void_p] * 3 + [C.c_int] * 6 + [C.c_void_p]
        d.coinn_conv3d_wgrad.argtypes = [C.c_void_p] * 3 + [C.c_int] * 6 + [C.c_void_p]
        d.coinn_conv3d_wgrad_tap.argtypes = [C.c_void_p] * 3 + [C.c_int] * 6 + [C.c_void_p]
        d.coinn_bn_pool_bwd_stats_pooled.argtypes = [C.c_void_p] * 5 + [C.c_longlong, C.c_int, C.c_void_p]
        d.coinn_pack_conv_weights.argtypes = [C.c_void_p] * 3 + [C.c_int] * 4 + [C.c_void_p]
        d.coinn_linear_small_fwd.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_int] * 4 + [C.c_void_p]
        d.coinn_linear_bn_small_fwd.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 10 + [C.c_int] * 5 + [C.c_float] * 2 + [C.c_void_p]
        d.coinn_linear_bn_small_bwd.argtypes = [C.c_void_p] * 6 + [C.c_int] + [C.c_void_p] * 6 + [C.c_int] * 4 + [C.c_void_p]
        d.coinn_linear_small_bwd.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.c_int] * 3 + [C.c_void_p]

        # MX-FP8 (csrc/mxfp8.cu)
        d.coinn_quantize_mx.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p]
        d.coinn_gemm_mxfp8_tn.argtypes = [C.c_void_p] * 6 + [C.c_int] * 8 + [C.c_void_p]
        d.coinn_quantize_mx_grouped.argtypes = [C.c_void_p, C.c_int, C.c_longlong, C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p]
        d.coinn_conv3d_mxfp8.argtypes = [C.c_void_p] * 5 + [C.c_int] * 6 + [C.c_void_p]
        # low-rank engines (csrc/lowrank.cu)
        d.coinn_psgd_mq.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        d.coinn_psgd_mtp.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        d.coinn_psgd_reconstruct.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        d.coinn_orthogonalize_batched.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p]
        d.coinn_segcopy.argtypes = [C.c_void_p, C.c_int, C.c_longlong, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
        d.coinn_gram_seg.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        d.coinn_lowrank_eig.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        d.coinn_skinny_gemm_seg.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_void_p]
        d.coinn_dad_reconstruct.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
        d.coinn_allgather.argtypes = [C.POINTER(AllGatherArgs), C.c_void_p]
        if d.coinn_allgather_args_size() != C.sizeof(AllGatherArgs):
            raise RuntimeError('AllGatherArgs ABI mismatch')

    def __getattr__(self, name):
        return getattr(self.dll, name)


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f'{LIB_PATH} is not built; run `python -m coinstac_dinunet_b200.ops.build`')
        _lib = _Lib()
    return _lib


def stream_ptr(device=None):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def check(code, what):
    if code != 0:
        raise RuntimeError(f'{what} failed with CUDA error {code}')


FLOAT_CODES = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}
INT_CODES = {torch.int64: 0, torch.int32: 1, torch.uint8: 2, torch.float32: 3}
