"""ctypes binding of libeva_sm100.so (C ABI declared in include/eva_sm100.h).

Raw device pointers, sizes and the current CUDA stream go across; no torch types.  There is no CPU
or eager fallback: if the library is missing or a tensor is not on a CUDA device the call raises.
"""
import ctypes
import os
import threading
import weakref

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), 'lib', 'libeva_sm100.so')
# development only: A/B a second build of the same ABI (tools/core_bench.py); never a different code path
LIB_PATH = os.environ.get('EVA_SM100_LIB', LIB_PATH)

EVA_F32, EVA_F16, EVA_BF16 = 0, 1, 2
_DTYPES = {torch.float32: EVA_F32, torch.float16: EVA_F16, torch.bfloat16: EVA_BF16}
LARA_MIS = {'mis-opt': 0, 'mis-bh': 1, 'mis-biased': 2}
LARA_SAMPLE_SINGLE, LARA_SAMPLE_ANTITHETIC, LARA_SAMPLE_MULTI = 0, 1, 2


class EvaHeadsView(ctypes.Structure):
    _fields_ = [('ptr', ctypes.c_void_p), ('stride_b', ctypes.c_int64), ('stride_n', ctypes.c_int64),
                ('stride_h', ctypes.c_int64)]


class EvaGeometry(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        'batch', 'heads', 'tokens', 'head_dim', 'dims', 'grid_h', 'grid_w', 'window', 'ext', 'halo_left_only',
        'chunk', 'chunk_ext', 'causal', 'mask_queries', 'mask_is_neg_inf', 'io_dtype', 'bias_toeplitz', 'keep_stats')]


class EvaAdaptive(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ('w_q', 'b_q', 'ln_gain_q', 'ln_bias_q', 'w_k', 'b_k', 'ln_gain_k',
                                               'ln_bias_k')] + [('mu_coeff', ctypes.c_float), ('ln_eps', ctypes.c_float)]


class LaraGeometry(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        'batch', 'heads', 'tokens', 'head_dim', 'dims', 'grid_h', 'grid_w', 'landmarks', 'per_token_proj', 'mixed',
        'mis_type', 'sample_mode', 'zero_padded', 'io_dtype')] + [('alpha_coeff', ctypes.c_float)]


class RfaGeometry(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        'batch', 'heads', 'tokens', 'head_dim', 'method', 'proj_dim', 'nu', 'feat_dim', 'cos_weighting', 'io_dtype')]


class SbGeometry(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in (
        'batch', 'heads', 'tokens', 'head_dim', 'dims', 'grid_h', 'grid_w', 'window', 'proj_dim', 'io_dtype')]


class RaGeometry(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ('batch', 'heads', 'tokens', 'head_dim', 'mode', 'io_dtype')]


RFA_METHOD = {'favorp': 0, 'relu': 1, 'fourier': 2, 'dpfp': 3, 'relu-only': 4, 'sigmoid-only': 5, 'given': 6}

_lib = None
_lock = threading.Lock()


class EvaKernelError(RuntimeError):
    """A libeva_sm100 entry point returned a non-zero status."""


def load():
    """Load (once) and return the shared library.  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} not found: build it with `python efficient-attention_b200/build.py` '
                '(nvcc, sm_100a). There is no CPU / eager fallback for this package.')
        lib = ctypes.CDLL(LIB_PATH)
        P, I64, SZ = ctypes.c_void_p, ctypes.c_int64, ctypes.c_size_t
        G, V, A, LG = ctypes.POINTER(EvaGeometry), ctypes.POINTER(EvaHeadsView), ctypes.POINTER(EvaAdaptive), ctypes.POINTER(LaraGeometry)
        lib.eva_sm100_abi_version.restype = ctypes.c_int
        lib.eva_last_error.restype = ctypes.c_char_p
        lib.eva_num_chunks.argtypes = [G]
        lib.eva_chunk_stats.argtypes = [G, V, V, V, P, A, P, P, P, P]
        lib.eva_window_attention.argtypes = [G, V, V, V, P, P, P, P, I64, P, P]
        lib.eva_window_attention_lse.argtypes = [G, V, V, V, P, P, P, P, I64, P, P, ctypes.POINTER(ctypes.c_int32), P]
        lib.eva_forward_workspace_bytes.argtypes = [G, ctypes.POINTER(SZ)]
        lib.eva_forward.argtypes = [G, V, V, V, P, A, P, P, I64, P, P, SZ, ctypes.POINTER(ctypes.c_int32), P]
        lib.eva_backward.argtypes = [G, V, V, V, P, A, P, P, I64, P, P, P, P, P, P, P, P, P, P]
        lib.lara_backward_step.argtypes = [ctypes.c_int32, ctypes.c_int32, P, P, P, P, P, P, P, P, P, P, P, I64, I64, ctypes.c_int32, ctypes.c_int32,
                                           ctypes.c_int32, ctypes.c_float, ctypes.c_float, P]
        lib.lara_forward_workspace_bytes.argtypes = [LG, ctypes.POINTER(SZ)]
        lib.lara_forward.argtypes = [LG, V, V, V, P, A, P, P, P, SZ, P]
        lib.lara_forward_given_landmarks.argtypes = [LG, V, V, V, P, P, P, P, P, SZ, P]
        RG, SG, AG = ctypes.POINTER(RfaGeometry), ctypes.POINTER(SbGeometry), ctypes.POINTER(RaGeometry)
        lib.rfa_feature_dim.argtypes = [RG]
        lib.rfa_forward_workspace_bytes.argtypes = [RG, ctypes.POINTER(SZ)]
        lib.rfa_forward.argtypes = [RG, V, V, V, P, P, P, P, P, P, SZ, P]
        lib.scatterbrain_forward_workspace_bytes.argtypes = [SG, ctypes.POINTER(SZ)]
        lib.scatterbrain_forward.argtypes = [SG, V, V, V, P, P, P, P, P, SZ, P]
        lib.ra_forward.argtypes = [AG, V, V, V, P, P, P, P, P, SZ, P]
        lib.ra_forward_workspace_bytes.argtypes = [AG, ctypes.POINTER(SZ)]
        lib.ra_sample.argtypes = [AG, V, V, ctypes.c_uint64, P, P, P]
        for fn in ('rfa_feature_dim', 'rfa_forward_workspace_bytes', 'rfa_forward', 'scatterbrain_forward_workspace_bytes',
                   'scatterbrain_forward', 'ra_forward', 'ra_forward_workspace_bytes', 'ra_sample'):
            getattr(lib, fn).restype = ctypes.c_int
        for fn in ('eva_num_chunks', 'eva_chunk_stats', 'eva_window_attention', 'eva_forward_workspace_bytes',
                   'eva_forward', 'eva_backward', 'eva_window_attention_lse', 'lara_backward_step', 'lara_forward_workspace_bytes', 'lara_forward', 'lara_forward_given_landmarks'):
            getattr(lib, fn).restype = ctypes.c_int
        if lib.eva_sm100_abi_version() != 4:
            raise RuntimeError('libeva_sm100.so ABI version mismatch; rebuild')
        _lib = lib
    return _lib


def _check(rc, what):
    if rc != 0:
        raise EvaKernelError(f'{what} failed ({rc}): {load().eva_last_error().decode()}')


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError('efficient_attention (B200 build) needs CUDA tensors: there is no CPU fallback')


def io_dtype(t):
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise TypeError(f'unsupported activation dtype {t.dtype} (float32, float16, bfloat16)') from None


def heads_view(t):
    """t: [B, N, H, D] view (any strides, D contiguous) -> EvaHeadsView."""
    assert t.dim() == 4 and t.stride(3) == 1, 'expected a [B, N, H, D] view with contiguous D'
    return EvaHeadsView(t.data_ptr(), t.stride(0), t.stride(1), t.stride(2))


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _f32(t):
    return None if t is None else t.detach().to(torch.float32).contiguous()


# module -> {tag: (key, value)}; weak keys: the cache dies with the module and stays out of its __dict__ (the values hold ctypes
# structures with raw pointers, which copy.deepcopy / pickle of a module -- timm's ModelEma, torch.save(model) -- must never see)
_MEMO = weakref.WeakKeyDictionary()


def memo(owner, tag, sources, build):
    """Per-module cache of tensors derived from parameters (fp32 copies, gathered bias tables, fused weights).

    An entry is rebuilt whenever a source tensor was modified in place through the parameter itself (optimizer step, `p.copy_`,
    load_state_dict: `_version` changes) or replaced (`.to()`, `.half()`: storage / dtype / device change).  Writes through
    `p.data` (`p.data.copy_(...)`: fairseq's FP16 optimizer sync, some EMA / weight-conversion scripts) do NOT move `_version`,
    so the key cannot see them; two rules close that hole:
      * a module in training mode never uses the cache and drops its entry (weights change every step anyway), so whatever
        a training loop does to `.data` is picked up by the next forward and by the first eval forward after it;
      * `invalidate_caches(module)` (re-exported by the package) is the explicit hook for scripts that write `.data` on a
        module that stays in eval mode.
    float32 contiguous sources are never copied at all (`_f32` returns a view of the parameter's own storage), so for float32
    modules only gathered tables (relative-position bias) are real copies.  The cache is keyed weakly by the module and dies
    with it.  Saves a dozen tiny conversion kernels and their host-side launches on every eval forward.
    """
    cache = _MEMO.get(owner)
    if getattr(owner, 'training', False):
        if cache is not None:
            cache.pop(tag, None)
        return build()
    key = tuple(None if t is None else (t.data_ptr(), t._version, t.dtype, t.device, tuple(t.shape)) for t in sources)
    if cache is None:
        cache = _MEMO[owner] = {}
    hit = cache.get(tag)
    if hit is not None and hit[0] == key:
        return hit[1]
    val = build()
    cache[tag] = (key, val)
    return val


def invalidate_caches(module=None):
    """Drop the parameter-derived caches of `module` and its sub-modules (all modules when None).  Needed only after writing
    parameters through `.data` on a module that stays in eval mode (see `memo`)."""
    if module is None:
        _MEMO.clear()
        return
    mods = module.modules() if hasattr(module, 'modules') else (module,)
    for m in mods:
        _MEMO.pop(m, None)


def adaptive(wq, bq, gq, betq, wk, bk, gk, betk, mu_coeff, ln_eps=1e-5):
    """Returns (struct, keepalive list of float32 tensors)."""
    keep = [_f32(t) for t in (wq, bq, gq, betq, wk, bk, gk, betk)]
    s = EvaAdaptive(*[None if t is None else t.data_ptr() for t in keep], mu_coeff, ln_eps)
    return s, keep


def _mask_u8(mask, B, N):
    if mask is None:
        return None
    m = mask.reshape(B, N).to(torch.bool).to(torch.uint8).contiguous()
    return m


def _stream(dev):
    return ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def eva_geometry(q, *, seq_shape, window, ext, chunk, chunk_ext, causal=False, halo_left_only=False,
                 mask_queries=False, mask_is_neg_inf=False, bias_toeplitz=False, keep_stats=False):
    B, N, H, D = q.shape
    two_d = len(seq_shape) == 2
    return EvaGeometry(B, H, N, D, 2 if two_d else 1, seq_shape[0] if two_d else 1, seq_shape[1] if two_d else N,
                       window, ext, int(halo_left_only), chunk, chunk_ext, int(causal), int(mask_queries),
                       int(mask_is_neg_inf), io_dtype(q), int(bias_toeplitz), int(keep_stats))


def num_chunks(geom):
    n = load().eva_num_chunks(ctypes.byref(geom))
    if n < 0:
        _check(n, 'eva_num_chunks')
    return n


def eva_forward(q, k, v, geom, ada, *, pad_mask=None, noise=None, bias=None, return_path=False, return_stats=False):
    """q,k,v: [B,N,H,D] views.  Returns out [B,N,H*D] (same dtype).  return_stats (geom.keep_stats must be set): also the forward's
    chunk statistics (k_bar, beta), float32 [B, H, C, D] views of the call's workspace."""
    lib = load()
    _require_cuda(q, k, v, pad_mask, noise, bias)
    B, N, H, D = q.shape
    ada_s, keep = ada
    mask = _mask_u8(pad_mask, B, N)
    noise = _f32(noise)
    bias = _f32(bias)
    out = torch.empty(B, N, H * D, dtype=q.dtype, device=q.device)
    nbytes = ctypes.c_size_t(0)
    _check(lib.eva_forward_workspace_bytes(ctypes.byref(geom), ctypes.byref(nbytes)), 'eva_forward_workspace_bytes')
    ws = torch.empty(max(nbytes.value, 256), dtype=torch.uint8, device=q.device)
    path = ctypes.c_int32(-1)
    bias_sh = 0 if bias is None or bias.shape[0] == 1 else bias.shape[1] * bias.shape[2]
    with torch.cuda.device(q.device):
        rc = lib.eva_forward(ctypes.byref(geom), ctypes.byref(heads_view(q)), ctypes.byref(heads_view(k)),
                             ctypes.byref(heads_view(v)), _ptr(mask), ctypes.byref(ada_s), _ptr(noise), _ptr(bias),
                             bias_sh, _ptr(out), _ptr(ws), ws.numel(), ctypes.byref(path), _stream(q.device))
    _check(rc, 'eva_forward')
    del keep
    path_id = path.value & 0xff
    if return_stats:
        assert geom.keep_stats, 'return_stats needs a geometry built with keep_stats=True'
        C = num_chunks(geom)
        n = B * H * C * D
        second = (n * 4 + 255) // 256 * 256
        stats = [ws[:n * 4].view(torch.float32).view(B, H, C, D), ws[second:second + n * 4].view(torch.float32).view(B, H, C, D)]
        if path.value & 0x100:               # the row log-sum-exp was kept too: the last region of the workspace
            nl = B * H * N * 4
            off = nbytes.value - (nl + 255) // 256 * 256
            stats.append(ws[off:off + nl].view(torch.float32).view(B, H, N))
        stats = tuple(stats)
        return (out, path_id, stats) if return_path else (out, stats)
    return (out, path_id) if return_path else out


def eva_backward(q, k, v, geom, ada, out, grad_out, *, pad_mask=None, noise=None, bias=None, want_bias_grad=False, stats=None,
                 packed_out=False, lse=None):
    """Gradients of eva_forward / eva_window_attention (`ada` None for chunk-less geometries).
    Returns (grad_qkv float32 [3, B, N, H, D], grad_bias float32 like bias or None, chunk_rows float32 [12, B, H, C, D] or None --
    the slots are listed at eva_backward in include/eva_sm100.h).  packed_out: grad_qkv is returned in q's dtype in the packed
    [B, N, 3, H, D] layout instead (written by the kernels themselves)."""
    lib = load()
    _require_cuda(q, k, v, out, grad_out, pad_mask, noise, bias)
    B, N, H, D = q.shape
    C = num_chunks(geom) if geom.chunk > 0 else 0
    k_bar, beta = stats[:2] if stats else (None, None)
    if lse is None and stats and len(stats) > 2:
        lse = stats[2]
    assert lse is None or (lse.dtype == torch.float32 and lse.is_contiguous())
    assert k_bar is None or (k_bar.dtype == torch.float32 and k_bar.is_contiguous() and beta.is_contiguous())
    mask = _mask_u8(pad_mask, B, N)
    noise = _f32(noise)
    bias = _f32(bias)
    out = out.detach().contiguous()
    grad_out = grad_out.detach().to(out.dtype).contiguous()
    grad_qkv = torch.empty(3, B, N, H, D, dtype=torch.float32, device=q.device)
    grad_io = torch.empty(B, N, 3, H, D, dtype=q.dtype, device=q.device) if packed_out else None
    rows = torch.empty(12, B, H, C, D, dtype=torch.float32, device=q.device) if C > 0 else None
    grad_bias = torch.empty_like(bias) if (want_bias_grad and bias is not None) else None
    ada_s, keep = ada if ada is not None else (None, None)
    bias_sh = 0 if bias is None or bias.shape[0] == 1 else bias.shape[1] * bias.shape[2]
    with torch.cuda.device(q.device):
        rc = lib.eva_backward(ctypes.byref(geom), ctypes.byref(heads_view(q)), ctypes.byref(heads_view(k)),
                              ctypes.byref(heads_view(v)), _ptr(mask), None if ada_s is None else ctypes.byref(ada_s), _ptr(noise),
                              _ptr(bias), bias_sh, _ptr(out), _ptr(grad_out), _ptr(k_bar), _ptr(beta), _ptr(lse), _ptr(grad_qkv), _ptr(grad_io),
                              _ptr(grad_bias), _ptr(rows), _stream(q.device))
    _check(rc, 'eva_backward')
    del keep
    return (grad_io if packed_out else grad_qkv), grad_bias, rows


def eva_chunk_stats(q, k, v, geom, ada, *, pad_mask=None, noise=None):
    lib = load()
    _require_cuda(q, k, v, pad_mask, noise)
    B, N, H, D = q.shape
    C = num_chunks(geom)
    ada_s, keep = ada
    mask = _mask_u8(pad_mask, B, N)
    noise = _f32(noise)
    k_bar = torch.empty(B, H, C, D, dtype=torch.float32, device=q.device)
    beta = torch.empty_like(k_bar)
    with torch.cuda.device(q.device):
        rc = lib.eva_chunk_stats(ctypes.byref(geom), ctypes.byref(heads_view(q)), ctypes.byref(heads_view(k)),
                                 ctypes.byref(heads_view(v)), _ptr(mask), ctypes.byref(ada_s), _ptr(noise),
                                 _ptr(k_bar), _ptr(beta), _stream(q.device))
    _check(rc, 'eva_chunk_stats')
    del keep
    return k_bar, beta


def eva_window_attention(q, k, v, geom, *, k_bar=None, beta=None, pad_mask=None, bias=None, return_lse=False):
    """return_lse: also the row log-sum-exp [B, H, N] float32 the backward can reuse, or None when the kernel that ran keeps none."""
    lib = load()
    _require_cuda(q, k, v, pad_mask, k_bar, beta, bias)
    B, N, H, D = q.shape
    mask = _mask_u8(pad_mask, B, N)
    bias = _f32(bias)
    k_bar, beta = _f32(k_bar), _f32(beta)
    out = torch.empty(B, N, H * D, dtype=q.dtype, device=q.device)
    bias_sh = 0 if bias is None or bias.shape[0] == 1 else bias.shape[1] * bias.shape[2]
    lse = torch.empty(B, H, N, dtype=torch.float32, device=q.device) if return_lse else None
    written = ctypes.c_int32(0)
    with torch.cuda.device(q.device):
        rc = lib.eva_window_attention_lse(ctypes.byref(geom), ctypes.byref(heads_view(q)), ctypes.byref(heads_view(k)),
                                          ctypes.byref(heads_view(v)), _ptr(mask), _ptr(k_bar), _ptr(beta), _ptr(bias),
                                          bias_sh, _ptr(out), _ptr(lse), ctypes.byref(written), _stream(q.device))
    _check(rc, 'eva_window_attention')
    if return_lse:
        return out, (lse if written.value else None)
    return out


def lara_backward_step(which, X, Y=None, *, dW=None, M2=None, v0=None, v1=None, v2=None, v3=None, o0=None, o1=None, o2=None,
                       x_item_stride=0, y_item_stride=0, items, landmarks, tokens, scale, alpha_coeff=0.0):
    """One of the three fused steps of the LARA backward (see lara_backward_step in include/eva_sm100.h)."""
    lib = load()
    _require_cuda(X, Y, dW, M2, v0, v1, v2, v3, o0, o1, o2)
    with torch.cuda.device(X.device):
        rc = lib.lara_backward_step(which, io_dtype(X), _ptr(X), _ptr(Y), _ptr(dW), _ptr(M2), _ptr(v0), _ptr(v1), _ptr(v2), _ptr(v3),
                                    _ptr(o0), _ptr(o1), _ptr(o2), x_item_stride, y_item_stride, items, landmarks, tokens, scale,
                                    alpha_coeff, _stream(X.device))
    _check(rc, 'lara_backward_step')


def lara_forward(q, k, v, *, seq_shape, landmarks, per_token_proj, mixed, mis_type, sample_mode, zero_padded,
                 alpha_coeff, proj, pad_mask=None, noise=None, given_landmarks=None):
    """`given_landmarks`: float32 [B, H, 3, C, D] (q_bar | k_bar before mixing | v_bar) computed by the caller
    (pool_module_type == 'dense'); `proj` is ignored then."""
    lib = load()
    _require_cuda(q, k, v, pad_mask, noise, given_landmarks)
    B, N, H, D = q.shape
    two_d = len(seq_shape) == 2
    geom = LaraGeometry(B, H, N, D, 2 if two_d else 1, seq_shape[0] if two_d else 1, seq_shape[1] if two_d else N,
                        landmarks, int(per_token_proj), int(mixed), LARA_MIS[mis_type], sample_mode, int(zero_padded),
                        io_dtype(q), float(alpha_coeff))
    proj_s, keep = proj
    mask = _mask_u8(pad_mask, B, N)
    noise = _f32(noise)
    out = torch.empty(B, N, H * D, dtype=q.dtype, device=q.device)
    nbytes = ctypes.c_size_t(0)
    _check(lib.lara_forward_workspace_bytes(ctypes.byref(geom), ctypes.byref(nbytes)), 'lara_forward_workspace_bytes')
    ws = torch.empty(max(nbytes.value, 256), dtype=torch.uint8, device=q.device)
    with torch.cuda.device(q.device):
        if given_landmarks is not None:
            given = _f32(given_landmarks)
            assert tuple(given.shape) == (B, H, 3, landmarks, D), 'given_landmarks must be [B, H, 3, C, D]'
            rc = lib.lara_forward_given_landmarks(ctypes.byref(geom), ctypes.byref(heads_view(q)), ctypes.byref(heads_view(k)),
                                                  ctypes.byref(heads_view(v)), _ptr(mask), _ptr(given), _ptr(noise), _ptr(out),
                                                  _ptr(ws), ws.numel(), _stream(q.device))
        else:
            rc = lib.lara_forward(ctypes.byref(geom), ctypes.byref(heads_view(q)), ctypes.byref(heads_view(k)),
                                  ctypes.byref(heads_view(v)), _ptr(mask), ctypes.byref(proj_s), _ptr(noise), _ptr(out),
                                  _ptr(ws), ws.numel(), _stream(q.device))
    _check(rc, 'lara_forward')
    del keep
    return out


def rfa_tc_launches():
    """How many times the tcgen05 Performer kernel has been launched by this process (diagnostic; tests assert the path)."""
    return load().eva_debug_rfa_tc_launches()


def rfa_forward(q, k, v, *, method, proj=None, nu=1, cos_weighting=False, pad_mask=None, q_feat=None, k_feat=None):
    """Linear attention with the feature map `method` (kernelized_attention.py:301-320).  q, k, v: [B, N, H, D] views; proj float32
    [H, m, D]; method 'given': q_feat / k_feat float32 [B, H, N, M] computed by the caller.  Returns [B, N, H*D] in q's dtype."""
    lib = load()
    _require_cuda(q, k, v, pad_mask, proj, q_feat, k_feat)
    B, N, H, D = q.shape
    proj, q_feat, k_feat = _f32(proj), _f32(q_feat), _f32(k_feat)
    geom = RfaGeometry(B, H, N, D, RFA_METHOD[method], 0 if proj is None else proj.shape[1], nu,
                       0 if q_feat is None else q_feat.shape[-1], int(cos_weighting), io_dtype(q))
    mask = _mask_u8(pad_mask, B, N)
    out = torch.empty(B, N, H * D, dtype=q.dtype, device=q.device)
    nbytes = ctypes.c_size_t(0)
    _check(lib.rfa_forward_workspace_bytes(ctypes.byref(geom), ctypes.byref(nbytes)), 'rfa_forward_workspace_bytes')
    ws = torch.empty(max(nbytes.value, 256), dtype=torch.uint8, device=q.device)
    with torch.cuda.device(q.device):
        rc = lib.rfa_forward(ctypes.byref(geom), ctypes.byref(heads_view(q)), ctypes.byref(heads_view(k)), ctypes.byref(heads_view(v)),
                             _ptr(mask), _ptr(proj), _ptr(q_feat), _ptr(k_feat), _ptr(out), _ptr(ws), ws.numel(), _stream(q.device))
    _check(rc, 'rfa_forward')
    return out


def scatterbrain_forward(q, k, v, *, seq_shape, window, proj, pad_mask=None, bias=None):
    """ScatterBrain core (scatterbrain_attention.py:95-160): halo-free windows + random-feature keys for the rest of the sequence."""
    lib = load()
    _require_cuda(q, k, v, pad_mask, proj, bias)
    B, N, H, D = q.shape
    proj, bias = _f32(proj), _f32(bias)
    two_d = len(seq_shape) == 2
    geom = SbGeometry(B, H, N, D, 2 if two_d else 1, seq_shape[0] if two_d else 1, seq_shape[1] if two_d else N, window,
                      proj.shape[1], io_dtype(q))
    mask = _mask_u8(pad_mask, B, N)
    out = torch.empty(B, N, H * D, dtype=q.dtype, device=q.device)
    nbytes = ctypes.c_size_t(0)
    _check(lib.scatterbrain_forward_workspace_bytes(ctypes.byref(geom), ctypes.byref(nbytes)), 'scatterbrain_forward_workspace_bytes')
    ws = torch.empty(max(nbytes.value, 256), dtype=torch.uint8, device=q.device)
    with torch.cuda.device(q.device):
        rc = lib.scatterbrain_forward(ctypes.byref(geom), ctypes.byref(heads_view(q)), ctypes.byref(heads_view(k)),
                                      ctypes.byref(heads_view(v)), _ptr(mask), _ptr(proj), _ptr(bias), _ptr(out), _ptr(ws), ws.numel(),
                                      _stream(q.device))
    _check(rc, 'scatterbrain_forward')
    return out


def ra_forward(q, k, v, *, mode, extra=None, k_ind=None, noise=None):
    """Randomized attention core (randomized_attention.py:24-55).  mode 'mean' | 'given' (extra: [B, N, H*D] in q's dtype) |
    'gather' (k_ind int64 [B, H, N]); noise float32 [B, H, N, D] or None."""
    lib = load()
    _require_cuda(q, k, v, extra, k_ind, noise)
    B, N, H, D = q.shape
    geom = RaGeometry(B, H, N, D, {'mean': 0, 'given': 1, 'gather': 2}[mode], io_dtype(q))
    noise = _f32(noise)
    if extra is not None:
        extra = extra.detach().to(q.dtype).contiguous()
    if k_ind is not None:
        k_ind = k_ind.to(torch.int64).contiguous()
    out = torch.empty(B, N, H * D, dtype=q.dtype, device=q.device)
    nbytes = ctypes.c_size_t(0)
    _check(lib.ra_forward_workspace_bytes(ctypes.byref(geom), ctypes.byref(nbytes)), 'ra_forward_workspace_bytes')
    ws = torch.empty(max(nbytes.value, 256), dtype=torch.uint8, device=q.device)
    with torch.cuda.device(q.device):
        rc = lib.ra_forward(ctypes.byref(geom), ctypes.byref(heads_view(q)), ctypes.byref(heads_view(k)), ctypes.byref(heads_view(v)),
                            _ptr(extra), _ptr(k_ind), _ptr(noise), _ptr(out), _ptr(ws), ws.numel(), _stream(q.device))
    _check(rc, 'ra_forward')
    return out


def ra_sample_supported(q):
    return q.shape[-1] == 64 and q.dtype in (torch.float16, torch.bfloat16)


def ra_sample(q, k, *, seed=0, gumbel=None):
    """One key index per query drawn from softmax(scale q k^T) by Gumbel-max (`ra_sample` in include/eva_sm100.h): int64 [B, H, N].
    `gumbel`: explicit float32 [B, H, N, N] noise instead of the seeded hash (tests)."""
    lib = load()
    _require_cuda(q, k, gumbel)
    B, N, H, D = q.shape
    geom = RaGeometry(B, H, N, D, 2, io_dtype(q))
    gumbel = _f32(gumbel)
    k_ind = torch.empty(B, H, N, dtype=torch.int64, device=q.device)
    with torch.cuda.device(q.device):
        rc = lib.ra_sample(ctypes.byref(geom), ctypes.byref(heads_view(q)), ctypes.byref(heads_view(k)), ctypes.c_uint64(seed & (2 ** 64 - 1)),
                           _ptr(gumbel), _ptr(k_ind), _stream(q.device))
    _check(rc, 'ra_sample')
    return k_ind
