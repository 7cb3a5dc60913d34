"""Torch-tensor front ends of the C-ABI compute entry points (include/copo_b200.h).  Every function enqueues
hand-written CUDA kernels of libcopo_b200.so on the current torch stream; torch only owns the memory.  There is no
CPU implementation: calling these without the library or without an sm_100 device raises."""
import ctypes

import torch

from . import _lib

c_int, c_float, c_size_t, c_u32, c_void_p = ctypes.c_int, ctypes.c_float, ctypes.c_size_t, ctypes.c_uint32, ctypes.c_void_p
P = _lib.ptr


def _f32(t):
    assert t.is_cuda and t.dtype == torch.float32, (t.device, t.dtype)
    return t


def _rows2d(t):
    """(pointer, rows, cols, row stride) of a float32 CUDA matrix whose last dim is contiguous."""
    _f32(t)
    assert t.dim() == 2 and t.stride(1) == 1, (t.shape, t.stride())
    return ctypes.c_void_p(t.data_ptr()), t.shape[0], t.shape[1], t.stride(0)


class PpoHeadArgs(ctypes.Structure):
    _fields_ = ([(n, ctypes.c_void_p) for n in ("logits", "actions", "old_logp", "old_logits", "adv")] +
                [("v_cur", ctypes.c_void_p * 3), ("v_old", ctypes.c_void_p * 3), ("v_tgt", ctypes.c_void_p * 3),
                 ("dlogits", ctypes.c_void_p), ("dv", ctypes.c_void_p * 3), ("stats", ctypes.c_void_p),
                 ("rows", ctypes.c_int32), ("n_heads", ctypes.c_int32), ("mode", ctypes.c_int32)] +
                [(n, ctypes.c_float) for n in ("clip_param", "vf_clip_param", "vf_loss_coeff", "entropy_coeff",
                                               "kl_coeff")] +
                [("norm_rows", ctypes.c_int32), ("plain_value_loss", ctypes.c_int32), ("dyn_coeffs", ctypes.c_void_p)])


class GaeArgs(ctypes.Structure):
    _fields_ = [("flags", ctypes.c_void_p), ("rewards", ctypes.c_void_p * 3), ("values", ctypes.c_void_p * 3),
                ("advantages", ctypes.c_void_p * 3), ("targets", ctypes.c_void_p * 3), ("T", ctypes.c_int32),
                ("N", ctypes.c_int32), ("heads", ctypes.c_int32), ("global_reward_per_scene", ctypes.c_int32),
                ("gamma", ctypes.c_float), ("lambda_", ctypes.c_float), ("bootstrap", ctypes.c_void_p * 3)]


def _lib_ready():
    return _lib.require_device()


def _dp(t):
    return t.data_ptr() if t is not None else None


# ---- linear layers -----------------------------------------------------------------------------------------------
def linear_forward(x, W, b, act, out=None):
    """act(x W^T + b); x [M, K], W [N, K] (torch Linear layout)."""
    lib = _lib_ready()
    px, M, K, ldx = _rows2d(x)
    N = W.shape[0]
    assert W.shape == (N, K) and W.is_contiguous()
    y = out if out is not None else torch.empty((M, N), dtype=torch.float32, device=x.device)
    py, _, _, ldy = _rows2d(y)
    if N <= 8:
        assert act == 0
        _lib.check(lib.b2c_head_forward(px, c_int(ldx), P(_f32(W)), P(b), py, c_int(ldy), c_int(M), c_int(K), c_int(N),
                                        _lib.stream_ptr()))
    else:
        _lib.check(lib.b2c_linear_forward(px, c_int(ldx), P(_f32(W)), P(b), py, c_int(ldy), c_int(M), c_int(K),
                                          c_int(N), c_int(act), _lib.stream_ptr()))
    return y


def head_backward_split(dy, h, W, dW, db, db_hidden=None, want_f32=True):
    """Backward of a narrow output layer on tanh activations h: returns (dz fp32 or None, dz as [hi | lo] bf16).
    db_hidden [K]: += column sums of dz, the bias gradient of the layer that produced h (then nobody needs dz fp32)."""
    lib = _lib_ready()
    pdy, M, N, ldy = _rows2d(dy)
    ph, _, K, ldh = _rows2d(h)
    dz = torch.empty((M, K), dtype=torch.float32, device=dy.device) if want_f32 else None
    dzs = torch.empty((M, 2 * K), dtype=torch.bfloat16, device=dy.device)
    _lib.check(lib.b2c_head_backward_tc(pdy, c_int(ldy), ph, c_int(ldh), P(W), P(dz), c_int(K), P(dzs), P(dW), P(db),
                                        P(db_hidden), c_int(M), c_int(K), c_int(N), c_int(1), _lib.stream_ptr()))
    return dz, dzs


def linear_backward(dy, x, W, dW, db, h_prev_is_tanh, need_dx=True, dx_out=None):
    """Backward of y = x W^T + b given dy.  Accumulates dW / db; returns dz_prev = (dy W) * (1 - x^2) when x is a tanh
    output (h_prev_is_tanh), else dy W; None when need_dx is False."""
    lib = _lib_ready()
    pdy, M, N, ldy = _rows2d(dy)
    px, _, K, ldx = _rows2d(x)
    assert W.shape == (N, K)
    dx = None
    if need_dx:
        dx = dx_out if dx_out is not None else torch.empty((M, K), dtype=torch.float32, device=dy.device)
    if N <= 8:
        pdx = P(dx)
        _lib.check(lib.b2c_head_backward(pdy, c_int(ldy), px, c_int(ldx), P(W), pdx, c_int(dx.stride(0) if need_dx else 0),
                                         P(dW), P(db), c_int(M), c_int(K), c_int(N), c_int(1 if h_prev_is_tanh else 0),
                                         _lib.stream_ptr()))
        return dx
    _lib.check(lib.b2c_linear_backward_weight(pdy, c_int(ldy), px, c_int(ldx), P(dW), P(db), c_int(M), c_int(K), c_int(N),
                                              _lib.stream_ptr()))
    if need_dx:
        _lib.check(lib.b2c_linear_backward_input(pdy, c_int(ldy), P(W), px if h_prev_is_tanh else None, c_int(ldx), P(dx),
                                                 c_int(dx.stride(0)), c_int(M), c_int(K), c_int(N), _lib.stream_ptr()))
    return dx


def gaussian_sample(logits, eps=None, seed=0, step=0, deterministic=False, want_eps=False):
    lib = _lib_ready()
    M = logits.shape[0]
    assert logits.shape == (M, 4) and logits.is_contiguous()
    actions = torch.empty((M, 2), dtype=torch.float32, device=logits.device)
    logp = torch.empty((M,), dtype=torch.float32, device=logits.device)
    eps_out = torch.empty((M, 2), dtype=torch.float32, device=logits.device) if want_eps else None
    _lib.check(lib.b2c_gaussian_sample(P(_f32(logits)), P(eps), P(actions), P(logp), P(eps_out), c_int(M),
                                       c_u32(seed & 0xFFFFFFFF), c_u32(step & 0xFFFFFFFF), c_int(int(deterministic)),
                                       _lib.stream_ptr()))
    return (actions, logp, eps_out) if want_eps else (actions, logp)


def ppo_head(logits, actions, old_logp, old_logits, adv, heads, cfg, mode=0, stats=None, norm_rows=0):
    """heads: list of (v_cur, v_old, v_tgt) triples.  Returns (dlogits, [dv...], stats[8] float64).
    norm_rows: rows the loss means are taken over (0: this call's rows; data parallel: the global minibatch size)."""
    lib = _lib_ready()
    M = logits.shape[0]
    dev = logits.device
    dlogits = torch.empty((M, 4), dtype=torch.float32, device=dev)
    dvs = [torch.empty((M,), dtype=torch.float32, device=dev) for _ in heads]
    if stats is None:
        stats = torch.zeros(8, dtype=torch.float64, device=dev)
    a = PpoHeadArgs()
    a.logits, a.actions = _dp(logits), _dp(actions)
    a.old_logp, a.old_logits, a.adv = _dp(old_logp), _dp(old_logits), _dp(adv)
    for k, (vc, vo, vt) in enumerate(heads):
        for t in (vc, vo, vt):
            assert t.is_contiguous() and t.dtype == torch.float32 and t.numel() == M
        a.v_cur[k], a.v_old[k], a.v_tgt[k], a.dv[k] = vc.data_ptr(), vo.data_ptr(), vt.data_ptr(), dvs[k].data_ptr()
    a.dlogits, a.stats = dlogits.data_ptr(), stats.data_ptr()
    a.rows, a.n_heads, a.mode = M, len(heads), mode
    a.clip_param, a.vf_clip_param = cfg["clip_param"], cfg["vf_clip_param"]
    a.vf_loss_coeff, a.entropy_coeff, a.kl_coeff = cfg["vf_loss_coeff"], cfg["entropy_coeff"], cfg["kl_coeff"]
    a.norm_rows, a.plain_value_loss = int(norm_rows), 0 if cfg.get("old_value_loss", True) else 1
    dyn = cfg.get("dyn_coeffs")                      # device [kl_coeff, entropy_coeff], read by the kernel at run time
    a.dyn_coeffs = dyn.data_ptr() if dyn is not None else None
    _lib.check(lib.b2c_ppo_head(ctypes.byref(a), _lib.stream_ptr()))
    return dlogits, dvs, stats


def lcf_meta_terms(adv, nei_adv, eps, lcf_mean=None, lcf_std=None, lcf_parameters=None):
    """Sums of the LCF side of the meta-gradient; the current LCF comes either as host floats or (no host read) from
    the model's raw `lcf_parameters` device tensor."""
    lib = _lib_ready()
    out = torch.zeros(3, dtype=torch.float64, device=adv.device)
    if lcf_parameters is not None:
        _lib.check(lib.b2c_lcf_meta_terms_params(P(_f32(adv)), P(_f32(nei_adv)), P(_f32(eps)), c_int(adv.numel()),
                                                 P(_f32(lcf_parameters)), P(out), _lib.stream_ptr()))
    else:
        _lib.check(lib.b2c_lcf_meta_terms(P(_f32(adv)), P(_f32(nei_adv)), P(_f32(eps)), c_int(adv.numel()),
                                          c_float(lcf_mean), c_float(lcf_std), P(out), _lib.stream_ptr()))
    return out


class LcfMetaFinishArgs(ctypes.Structure):              # b2c_lcf_meta_finish_args
    _fields_ = [("grad_value", c_void_p), ("st_new", c_void_p), ("st_old", c_void_p), ("sums", c_void_p),
                ("rows", ctypes.c_double), ("raw_mean", ctypes.c_double), ("raw_std", ctypes.c_double),
                ("lcf_parameters", c_void_p), ("exp_avg", c_void_p), ("exp_avg_sq", c_void_p), ("lcf_grad", c_void_p),
                ("stats", c_void_p), ("lr", c_float), ("beta1", c_float), ("beta2", c_float), ("eps", c_float),
                ("step", ctypes.c_int32)]


def lcf_meta_sums(adv, nei_adv, eps, global_adv, lcf_parameters, out4):
    """out4 (float64, zeroed by the caller) += sums of: coordinated advantage, d, d * eps (lcf_meta_terms) and the global
    advantage - one pass over the minibatch's columns."""
    lib = _lib_ready()
    _lib.check(lib.b2c_lcf_meta_sums(P(_f32(adv)), P(_f32(nei_adv)), P(_f32(eps)), P(_f32(global_adv)), c_int(adv.numel()),
                                     P(_f32(lcf_parameters)), P(out4), _lib.stream_ptr()))
    return out4


def lcf_meta_finish(grad_value, st_new, st_old, sums, rows, raw_mean, raw_std, lcf_parameters, exp_avg, exp_avg_sq,
                    lcf_grad, stats, lr, step, beta1=0.9, beta2=0.999, eps=1e-8):
    """The tail of CoPOPolicy.meta_update in one launch (b2c_lcf_meta_finish): LCF loss and gradient, Adam step on
    lcf_parameters, the 13 logged statistics into `stats` (float64)."""
    lib = _lib_ready()
    a = LcfMetaFinishArgs()
    for name, t, dt, n in (("grad_value", grad_value, torch.float64, 1), ("st_new", st_new, torch.float64, 8),
                           ("st_old", st_old, torch.float64, 8), ("sums", sums, torch.float64, 4),
                           ("lcf_parameters", lcf_parameters, torch.float32, 2), ("exp_avg", exp_avg, torch.float32, 2),
                           ("exp_avg_sq", exp_avg_sq, torch.float32, 2), ("lcf_grad", lcf_grad, torch.float32, 2),
                           ("stats", stats, torch.float64, 13)):
        assert t.dtype == dt and t.numel() >= n and t.is_contiguous(), name
        setattr(a, name, t.data_ptr())
    a.rows, a.raw_mean, a.raw_std = float(rows), float(raw_mean), float(raw_std)
    a.lr, a.beta1, a.beta2, a.eps, a.step = lr, beta1, beta2, eps, int(step)
    _lib.check(lib.b2c_lcf_meta_finish(ctypes.byref(a), _lib.stream_ptr()))
    return stats


# ---- rollout bookkeeping ------------------------------------------------------------------------------------------
def gae3(flags, rewards, values, gamma, lambda_, global_reward_per_scene=0, bootstrap=None):
    """flags uint8 [T, N]; rewards / values: lists of 1 or 3 float32 [T, N] tensors (the global reward may be [T, S]
    with global_reward_per_scene = slots per scene).  bootstrap: optional list of [N] tensors (or None) per head - the
    value of the observation after the last row, used for trajectories the fragment end cuts (stock rllib rule);
    without it such a trajectory bootstraps with the value of its own last row (the CCPPO / CoPO rule).
    Returns (advantages, targets) lists."""
    lib = _lib_ready()
    T, N = flags.shape
    heads = len(values)
    adv = [torch.empty((T, N), dtype=torch.float32, device=flags.device) for _ in range(heads)]
    tgt = [torch.empty((T, N), dtype=torch.float32, device=flags.device) for _ in range(heads)]
    a = GaeArgs()
    a.flags = flags.data_ptr()
    for h in range(heads):
        assert rewards[h].is_contiguous() and values[h].is_contiguous()
        a.rewards[h], a.values[h] = rewards[h].data_ptr(), values[h].data_ptr()
        a.advantages[h], a.targets[h] = adv[h].data_ptr(), tgt[h].data_ptr()
        if bootstrap is not None and bootstrap[h] is not None:
            b = bootstrap[h]
            assert b.dtype == torch.float32 and b.is_contiguous() and b.numel() == N
            a.bootstrap[h] = b.data_ptr()
    a.T, a.N, a.heads, a.global_reward_per_scene = T, N, heads, global_reward_per_scene
    a.gamma, a.lambda_ = gamma, lambda_
    _lib.check(lib.b2c_gae3(ctypes.byref(a), _lib.stream_ptr()))
    return adv, tgt


def lcf_mix_stats(flags, adv, nei_adv, step_lcf, global_adv, out=None):
    lib = _lib_ready()
    if out is None:
        out = torch.zeros(5, dtype=torch.float64, device=adv.device)
    _lib.check(lib.b2c_lcf_mix_stats(P(flags), P(adv), P(nei_adv), P(step_lcf), P(global_adv), c_size_t(adv.numel()),
                                     P(out), _lib.stream_ptr()))
    return out


def lcf_mix_apply(flags, adv, nei_adv, step_lcf, global_adv, mean, std, gmean, gstd):
    lib = _lib_ready()
    norm = torch.empty_like(adv)
    _lib.check(lib.b2c_lcf_mix_apply(P(flags), P(adv), P(nei_adv), P(step_lcf), P(global_adv), P(norm),
                                     c_size_t(adv.numel()), c_float(mean), c_float(std), c_float(gmean), c_float(gstd),
                                     _lib.stream_ptr()))
    return norm


def stats_to_mean_std(sum_, sumsq, n):
    mean = sum_ / max(n, 1.0)
    var = max(sumsq / max(n, 1.0) - mean * mean, 0.0)
    return mean, var ** 0.5


def cc_obs_fuse(obs, actions, flags, mf_mask, nei_list, slots, mode, counterfactual=True, want_split=False):
    """obs [R, D] with R ordered [t][scene][slot]; mode 'none' | 'mf' | 'concat'.  want_split (mean-field mode on whole
    scenes): also returns the rows as the value network's [hi | lo] tensor-core operand (what tc_split_rows(cobs) would
    give) -> (cobs, cobs_split)."""
    lib = _lib_ready()
    R, D = obs.shape
    AD = actions.shape[1] if actions is not None else 2
    m = {"none": 0, "mf": 1, "concat": 2}[mode]
    n_other = (0, 1, 4)[m]
    C = D + n_other * (D + (AD if counterfactual else 0))
    cobs = torch.empty((R, C), dtype=torch.float32, device=obs.device)
    if want_split:
        assert m == 1 and R % slots == 0 and slots <= 64
        kp = tc_padded_k(C)
        split = torch.empty((R, 2 * kp), dtype=torch.bfloat16, device=obs.device)
        _lib.check(lib.b2c_cc_obs_fuse_split(P(obs), P(actions), P(flags), P(mf_mask), P(nei_list), P(cobs), P(split), c_int(kp),
                                             c_size_t(R), c_int(slots), c_int(D), c_int(AD), c_int(C), c_int(m),
                                             c_int(int(counterfactual)), _lib.stream_ptr()))
        return cobs, split
    _lib.check(lib.b2c_cc_obs_fuse(P(obs), P(actions), P(flags), P(mf_mask), P(nei_list), P(cobs), c_size_t(R), c_int(slots),
                                   c_int(D), c_int(AD), c_int(C), c_int(m), c_int(int(counterfactual)), _lib.stream_ptr()))
    return cobs


def gather_rows(src, idx, out=None):
    lib = _lib_ready()
    src2 = src if src.dim() == 2 else src.unsqueeze(1)
    assert src2.stride(1) == 1 and idx.dtype == torch.int64 and idx.is_contiguous()
    R, Wd = idx.numel(), src2.shape[1]
    dst = out if out is not None else torch.empty((R, Wd), dtype=torch.float32, device=src.device)
    _lib.check(lib.b2c_gather_rows(P(src2) if src2.is_contiguous() else ctypes.c_void_p(src2.data_ptr()),
                                   c_size_t(src2.stride(0)), P(idx), ctypes.c_void_p(dst.data_ptr()),
                                   c_size_t(dst.stride(0)), c_size_t(R), c_int(Wd), _lib.stream_ptr()))
    return dst if src.dim() == 2 else dst.squeeze(1)


def gather_cols(src, idx):
    """src [R, W] row-major, idx int64 [B] -> [W, B]: column c of the minibatch is the contiguous vector out[c]."""
    lib = _lib_ready()
    assert src.dim() == 2 and src.stride(1) == 1 and idx.dtype == torch.int64 and idx.is_contiguous()
    B, W = idx.numel(), src.shape[1]
    out = torch.empty((W, B), dtype=torch.float32, device=src.device)
    _lib.check(lib.b2c_gather_cols(P(src) if src.is_contiguous() else ctypes.c_void_p(src.data_ptr()), c_size_t(src.stride(0)),
                                   P(idx), P(out), c_size_t(B), c_int(W), _lib.stream_ptr()))
    return out


def adam_step(param, grad, exp_avg, exp_avg_sq, lr, step, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0):
    lib = _lib_ready()
    for t in (param, grad, exp_avg, exp_avg_sq):
        assert t.is_contiguous() and t.dtype == torch.float32 and t.numel() == param.numel()
    _lib.check(lib.b2c_adam_step(P(param), P(grad), P(exp_avg), P(exp_avg_sq), c_size_t(param.numel()), c_float(lr),
                                 c_float(beta1), c_float(beta2), c_float(eps), c_int(step), c_float(grad_scale),
                                 _lib.stream_ptr()))


def dot(a, b, out=None):
    lib = _lib_ready()
    if out is None:
        out = torch.zeros(1, dtype=torch.float64, device=a.device)
    _lib.check(lib.b2c_dot(P(a), P(b), c_size_t(a.numel()), P(out), _lib.stream_ptr()))
    return out


# ---- tcgen05 path (split bf16) --------------------------------------------------------------------------------------
def tc_padded_k(K):
    return (K + 63) // 64 * 64


def tc_split_rows(x, out=None, ones_col=False):
    """fp32 [M, K] -> bf16 [M, 2*Kp] (hi | lo).  ones_col (needs K < Kp): 1.0 in the first padding column, so that
    tc_wgrad(..., db=) can return the bias gradient with the weight gradient."""
    lib = _lib_ready()
    px, M, K, ldx = _rows2d(x)
    Kp = tc_padded_k(K)
    if out is None:
        out = torch.empty((M, 2 * Kp), dtype=torch.bfloat16, device=x.device)
    _lib.check(lib.b2c_tc_split_rows_ones(px, c_int(ldx), P(out), c_int(M), c_int(K), c_int(Kp), c_int(int(ones_col)),
                                          _lib.stream_ptr()))
    return out


def tc_has_ones_col(K):
    """True when a [*, K] input has a padding column for the ones of tc_split_rows(ones_col=True)."""
    return K < tc_padded_k(K)


def tc_prep_weight(W, transpose=False):
    """fp32 W [N, K] -> bf16 [rows, 2*Kp] = [hi | lo] of W (or of W^T when transpose)."""
    lib = _lib_ready()
    N, K = W.shape
    rows, red = (K, N) if transpose else (N, K)
    Kp = tc_padded_k(red)
    out = torch.empty((rows, 2 * Kp), dtype=torch.bfloat16, device=W.device)
    _lib.check(lib.b2c_tc_prep_weight(P(_f32(W)), P(out), c_int(N), c_int(K), c_int(Kp), c_int(int(transpose)),
                                      _lib.stream_ptr()))
    return out


def tc_linear(a_split, w_prep, bias=None, act=0, want_f32=True, want_split=False, dtanh_src=None, out_f32=None,
              out_split=None):
    """256-wide layer on the tensor cores.  Returns (fp32 [M, 256] or None, bf16 split [M, 512] or None)."""
    lib = _lib_ready()
    M, two_kp = a_split.shape
    Kp = two_kp // 2
    assert a_split.dtype == torch.bfloat16 and a_split.is_contiguous()
    assert w_prep.shape == (256, 2 * Kp) and w_prep.is_contiguous(), (w_prep.shape, Kp)
    dev = a_split.device
    if want_f32 and out_f32 is None:
        out_f32 = torch.empty((M, 256), dtype=torch.float32, device=dev)
    if want_split and out_split is None:
        out_split = torch.empty((M, 512), dtype=torch.bfloat16, device=dev)
    ld_src = dtanh_src.stride(0) if dtanh_src is not None else 0
    _lib.check(lib.b2c_tc_linear(P(a_split), P(w_prep), P(bias), ctypes.c_void_p(dtanh_src.data_ptr()) if dtanh_src is not None else None,
                                 c_int(ld_src), ctypes.c_void_p(out_f32.data_ptr()) if out_f32 is not None else None,
                                 c_int(out_f32.stride(0) if out_f32 is not None else 0), P(out_split), c_int(M), c_int(Kp),
                                 c_int(act), _lib.stream_ptr()))
    return out_f32, out_split


_WGRAD_WS = {}


def tc_linear_dgrad(dz_split, wT_prep, h_split, want_f32=False):
    """Input gradient of a 256-wide tanh layer: (dz W) * (1 - h^2) with h = hi + lo taken from the layer's own operand
    h_split [M, 512].  Returns (fp32 [M, 256] or None, bf16 split [M, 512])."""
    lib = _lib_ready()
    M = dz_split.shape[0]
    Kp = dz_split.shape[1] // 2
    assert h_split.shape == (M, 512) and h_split.is_contiguous() and wT_prep.shape == (256, 2 * Kp)
    out_f32 = torch.empty((M, 256), dtype=torch.float32, device=dz_split.device) if want_f32 else None
    out_split = torch.empty((M, 512), dtype=torch.bfloat16, device=dz_split.device)
    _lib.check(lib.b2c_tc_linear_dgrad(P(dz_split), P(wT_prep), P(h_split), P(out_f32), c_int(256 if want_f32 else 0),
                                       P(out_split), c_int(M), c_int(Kp), _lib.stream_ptr()))
    return out_f32, out_split


def tc_wgrad(dz_split, x_split, dW, db=None):
    """dW[256, K] += dz^T x on the tensor cores from the [hi | lo] operands (dz_split [M, 512], x_split [M, 2*Kp]).
    db: x_split was made with ones_col=True; db[256] += sum of dz over the rows (column K of the product)."""
    lib = _lib_ready()
    M = dz_split.shape[0]
    Kp = x_split.shape[1] // 2
    K = dW.shape[1]
    assert dz_split.shape == (M, 512) and x_split.shape[0] == M and dW.shape[0] == 256 and dW.is_contiguous()
    key = (dz_split.device, Kp)
    if key not in _WGRAD_WS:
        parts = lib.b2c_tc_wgrad_parts()
        _WGRAD_WS[key] = torch.empty(parts * 256 * Kp, dtype=torch.float32, device=dz_split.device)
    _lib.check(lib.b2c_tc_wgrad_bias(P(dz_split), P(x_split), P(_WGRAD_WS[key]), P(dW), P(db), c_int(M), c_int(K),
                                     c_int(Kp), _lib.stream_ptr()))
    return dW


def colsum(dy, db):
    lib = _lib_ready()
    pdy, M, N, ldy = _rows2d(dy)
    _lib.check(lib.b2c_colsum(pdy, c_int(ldy), P(db), c_int(M), c_int(N), _lib.stream_ptr()))
    return db


class TcHead(ctypes.Structure):
    _fields_ = [("weight", ctypes.c_void_p), ("bias", ctypes.c_void_p), ("out", ctypes.c_void_p), ("n", ctypes.c_int32),
                ("actions", ctypes.c_void_p), ("logp", ctypes.c_void_p), ("seed", ctypes.c_uint32),
                ("step", ctypes.c_uint32)]


def tc_linear_head(a_split, w_prep, bias, head_w, head_b, act=1, sample=None, want_f32=False, out=None, actions=None,
                   logp=None):
    """256-wide layer + fused narrow output layer (n = 1 or 4).  sample = (seed, step) also draws the Gaussian action
    in the epilogue.  Returns (head_out [M, n], actions or None, logp or None, hidden fp32 or None)."""
    lib = _lib_ready()
    M, two_kp = a_split.shape
    Kp = two_kp // 2
    n = head_w.shape[0]
    dev = a_split.device
    assert head_w.shape == (n, 256) and head_w.is_contiguous() and w_prep.shape == (256, 2 * Kp)
    if out is None:
        out = torch.empty((M, n), dtype=torch.float32, device=dev)
    h = torch.empty((M, 256), dtype=torch.float32, device=dev) if want_f32 else None
    hd = TcHead()
    hd.weight, hd.bias, hd.out, hd.n = head_w.data_ptr(), head_b.data_ptr(), out.data_ptr(), n
    if sample is not None:
        if actions is None:
            actions = torch.empty((M, 2), dtype=torch.float32, device=dev)
        if logp is None:
            logp = torch.empty((M,), dtype=torch.float32, device=dev)
        hd.actions, hd.logp = actions.data_ptr(), logp.data_ptr()
        hd.seed, hd.step = sample[0] & 0xFFFFFFFF, sample[1] & 0xFFFFFFFF
    _lib.check(lib.b2c_tc_linear_head(P(a_split), P(w_prep), P(bias), P(h), c_int(256 if want_f32 else 0), c_int(M),
                                      c_int(Kp), c_int(act), ctypes.byref(hd), _lib.stream_ptr()))
    return out, actions, logp, h


def tc_mlp2_head(a_split, w1_prep, b1, w2_prep, b2, head_w, head_b, sample=None, out=None, actions=None, logp=None,
                 train=False):
    """The whole inference pass of a 256-256 tanh network in ONE kernel (mlp_fused.cu): both hidden layers stay on the
    SM, the narrow output layer (n = 1 or 4) and - with sample = (seed, step) - the Gaussian action draw ride in the last
    epilogue.  Bit-identical to tc_linear + tc_linear_head.  Returns (head_out [M, n], actions or None, logp or None).
    train=True (the learner's forward, b2c_tc_mlp2_train): also writes the hidden layers the backward pass reads again and
    returns (head_out, h1_split [M, 512] bf16, h2 [M, 256] fp32)."""
    lib = _lib_ready()
    M, two_kp = a_split.shape
    Kp = two_kp // 2
    n = head_w.shape[0]
    dev = a_split.device
    assert a_split.dtype == torch.bfloat16 and a_split.is_contiguous()
    assert w1_prep.shape == (256, 2 * Kp) and w2_prep.shape == (256, 512) and head_w.shape == (n, 256)
    assert w1_prep.is_contiguous() and w2_prep.is_contiguous() and head_w.is_contiguous()
    if out is None:
        out = torch.empty((M, n), dtype=torch.float32, device=dev)
    hd = TcHead()
    hd.weight, hd.bias, hd.out, hd.n = head_w.data_ptr(), head_b.data_ptr(), out.data_ptr(), n
    if sample is not None:
        if actions is None:
            actions = torch.empty((M, 2), dtype=torch.float32, device=dev)
        if logp is None:
            logp = torch.empty((M,), dtype=torch.float32, device=dev)
        hd.actions, hd.logp = actions.data_ptr(), logp.data_ptr()
        hd.seed, hd.step = sample[0] & 0xFFFFFFFF, sample[1] & 0xFFFFFFFF
    if train:
        h1_split = torch.empty((M, 512), dtype=torch.bfloat16, device=dev)
        h2 = torch.empty((M, 256), dtype=torch.float32, device=dev)
        _lib.check(lib.b2c_tc_mlp2_train(P(a_split), c_int(Kp), P(w1_prep), P(b1), P(w2_prep), P(b2), ctypes.byref(hd),
                                         P(h1_split), P(h2), c_int(M), _lib.stream_ptr()))
        return out, h1_split, h2
    _lib.check(lib.b2c_tc_mlp2_head(P(a_split), c_int(Kp), P(w1_prep), P(b1), P(w2_prep), P(b2), ctypes.byref(hd), c_int(M),
                                    _lib.stream_ptr()))
    return out, actions, logp
