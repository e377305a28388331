"""Input preparation of the TTT op on libttt_b200.so: mirror of ``TTTBase.process_input``
(reference: ttt/models/ssm/ttt_layer.py:252-306).  The q/k/v Linears and the lr projection stay library GEMMs; everything
between them and the scan (L2 norm, RoPE, reconstruction target, mini-batch transpose, interleave, eta) is one kernel
forward and one kernel backward (csrc/process_input.cu); there is no eager fallback."""
import torch
import torch.nn.functional as F

from . import _lib


def _prepare_fwd(xq, xk, xv, lr_logit, rope_cos, rope_sin, ln_w, ln_b, seq_text_length, mini_batch_size, ttt_base_lr, index=None):
    B, L, E = xq.shape
    H = E // 64
    if E != H * 64 or L % mini_batch_size:
        raise RuntimeError("process_input: model_dim must be heads x 64 and L a multiple of the mini-batch size")
    for t, n in ((xq, "xq"), (xk, "xk"), (xv, "xv")):
        if not (t.is_cuda and t.dtype == torch.bfloat16 and t.is_contiguous() and t.shape == xq.shape):
            raise RuntimeError(f"{n} must be a contiguous CUDA bf16 tensor [B, L, H*64]")
    dev = xq.device
    f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()
    lg, c, s, lw, lb = f32(lr_logit), f32(rope_cos), f32(rope_sin), f32(ln_w).reshape(H, 64), f32(ln_b).reshape(H, 64)
    if lg.shape != (B, L, H) or c.shape[-1] != 32 or c.shape[0] < L - seq_text_length:
        raise RuntimeError("process_input: lr_logit must be [B,L,H] and the RoPE tables [>= video tokens, 32]")
    idx = None if index is None else index.to(device=dev, dtype=torch.int32).contiguous()
    NC = L // mini_batch_size
    XQ = torch.empty(B, H, NC, mini_batch_size, 64, device=dev, dtype=torch.bfloat16)
    XK, XV = torch.empty_like(XQ), torch.empty_like(XQ)
    eta = torch.empty(B, H, NC, mini_batch_size, device=dev, dtype=torch.bfloat16)
    p = _lib.ptr
    code = _lib.lib().ttt_b200_process_input(p(xq), p(xk), p(xv), p(lg), p(c), p(s), p(lw), p(lb), p(idx), p(XQ), p(XK), p(XV),
                                             p(eta), B, L, H, int(seq_text_length), int(mini_batch_size), float(ttt_base_lr),
                                             _lib.current_stream(xq))
    _lib.check(code, "ttt_b200_process_input")
    return (XQ, XK, XV, eta), (lg, c, s, lw, idx)


class _Prepare(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xq, xk, xv, lr_logit, ln_w, ln_b, rope_cos, rope_sin, seq_text_length, mini_batch_size, ttt_base_lr, index):
        outs, (lg, c, s, lw, idx) = _prepare_fwd(xq, xk, xv, lr_logit, rope_cos, rope_sin, ln_w, ln_b, seq_text_length,
                                                 mini_batch_size, ttt_base_lr, index)
        ctx.save_for_backward(xq, xk, xv, lg, c, s, lw, *(() if idx is None else (idx,)))
        ctx.cfg = (int(seq_text_length), int(mini_batch_size), float(ttt_base_lr), idx is not None,
                   lr_logit.dtype, ln_w.dtype, ln_b.dtype, ln_w.shape, ln_b.shape)
        return outs

    @staticmethod
    def backward(ctx, gQ, gK, gV, gEta):
        xq, xk, xv, lg, c, s, lw, *rest = ctx.saved_tensors
        seq_text, CS, base_lr, has_idx, lg_dt, lw_dt, lb_dt, lw_shape, lb_shape = ctx.cfg
        idx = rest[0] if has_idx else None
        B, L, E = xq.shape
        H = E // 64
        bf = lambda t: t.to(torch.bfloat16).contiguous()
        gQ, gK, gV = bf(gQ), bf(gK), bf(gV)
        ge = gEta.float().contiguous()
        gxq, gxk, gxv = torch.empty_like(xq), torch.empty_like(xk), torch.empty_like(xv)
        glg = torch.empty(B, L, H, device=xq.device, dtype=torch.float32)
        glw = torch.empty(H, 64, device=xq.device, dtype=torch.float32)
        glb = torch.empty(H, 64, device=xq.device, dtype=torch.float32)
        p = _lib.ptr
        code = _lib.lib().ttt_b200_process_input_backward(p(xq), p(xk), p(xv), p(lg), p(c), p(s), p(lw), p(idx), p(gQ), p(gK),
                                                          p(gV), p(ge), p(gxq), p(gxk), p(gxv), p(glg), p(glw), p(glb), B, L, H,
                                                          seq_text, CS, base_lr, _lib.current_stream(xq))
        _lib.check(code, "ttt_b200_process_input_backward")
        return (gxq, gxk, gxv, glg.to(lg_dt), glw.reshape(lw_shape).to(lw_dt), glb.reshape(lb_shape).to(lb_dt),
                None, None, None, None, None, None)


def prepare(xq, xk, xv, lr_logit, rope_cos, rope_sin, ln_w, ln_b, seq_text_length, mini_batch_size, ttt_base_lr, index=None):
    """xq/xk/xv bf16 [B,L,H*64]; lr_logit [B,L,H]; rope_cos/sin [Lv,32]; index: int32 [L] gather index of the multi-scene
    interleave (None for one scene).  Returns XQ, XK, XV bf16 [B,H,NC,CS,64] and last_eta bf16 [B,H,NC,CS]; differentiable
    w.r.t. xq, xk, xv, lr_logit, ln_w, ln_b."""
    return _Prepare.apply(xq, xk, xv, lr_logit, ln_w, ln_b, rope_cos, rope_sin, seq_text_length, mini_batch_size, ttt_base_lr,
                          index)


def process_input(hidden_states, P, rope_cos, rope_sin, seq_text_length, mini_batch_size, ttt_base_lr, index=None):
    """``TTTBase.process_input`` (ttt_layer.py:252-306).  P: dict with wq/wk/wv ``.weight``/``.bias``,
    ``learnable_ttt_lr_weight`` [H,1,E], ``learnable_ttt_lr_bias`` [H,1], ``ttt_norm_weight``/``ttt_norm_bias`` [H,64].
    Returns {"XQ","XK","XV"} [B,H,NC,CS,64] and "last_eta" [B,H,NC,CS]."""
    xq = F.linear(hidden_states, P["wq.weight"], P["wq.bias"]).contiguous()
    xk = F.linear(hidden_states, P["wk.weight"], P["wk.bias"]).contiguous()
    xv = F.linear(hidden_states, P["wv.weight"], P["wv.bias"]).contiguous()
    H = P["learnable_ttt_lr_weight"].shape[0]
    logit = F.linear(hidden_states, P["learnable_ttt_lr_weight"].reshape(H, -1), P["learnable_ttt_lr_bias"].reshape(H))
    XQ, XK, XV, eta = prepare(xq, xk, xv, logit, rope_cos, rope_sin, P["ttt_norm_weight"], P["ttt_norm_bias"], seq_text_length,
                              mini_batch_size, ttt_base_lr, index)
    return {"XQ": XQ, "XK": XK, "XV": XV, "last_eta": eta}


class _OutputNorm(torch.autograd.Function):
    @staticmethod
    def forward(ctx, op_out, post_norm_weight, post_norm_bias, eps, undo_index):
        B, H, NC, CS, Fd = op_out.shape
        if Fd != 64 or not (op_out.is_cuda and op_out.dtype == torch.bfloat16 and op_out.is_contiguous()):
            raise RuntimeError("output_norm: op_out must be a contiguous CUDA bf16 tensor [B,H,NC,CS,64]")
        dev = op_out.device
        L = NC * CS
        g = post_norm_weight.detach().to(device=dev, dtype=torch.float32).contiguous()
        bt = post_norm_bias.detach().to(device=dev, dtype=torch.float32).contiguous()
        idx = None if undo_index is None else undo_index.to(device=dev, dtype=torch.int32).contiguous()
        out = torch.empty(B, L, H * 64, device=dev, dtype=torch.bfloat16)
        p = _lib.ptr
        code = _lib.lib().ttt_b200_output_norm(p(op_out), p(g), p(bt), p(idx), p(out), B, L, H, float(eps), _lib.current_stream(op_out))
        _lib.check(code, "ttt_b200_output_norm")
        ctx.save_for_backward(op_out, g, *(() if idx is None else (idx,)))
        ctx.cfg = (float(eps), idx is not None, post_norm_weight.dtype, post_norm_bias.dtype)
        return out

    @staticmethod
    def backward(ctx, gout):
        op_out, g, *rest = ctx.saved_tensors
        eps, has_idx, w_dt, b_dt = ctx.cfg
        idx = rest[0] if has_idx else None
        B, H, NC, CS, _ = op_out.shape
        gout = gout.to(torch.bfloat16).contiguous()
        gop = torch.empty_like(op_out)
        dg = torch.empty(H * 64, device=op_out.device, dtype=torch.float32)
        db = torch.empty(H * 64, device=op_out.device, dtype=torch.float32)
        p = _lib.ptr
        code = _lib.lib().ttt_b200_output_norm_backward(p(op_out), p(g), p(idx), p(gout), p(gop), p(dg), p(db), B, NC * CS, H, eps,
                                                        _lib.current_stream(op_out))
        _lib.check(code, "ttt_b200_output_norm_backward")
        return gop, dg.to(w_dt), db.to(b_dt), None, None


def output_norm(op_out, post_norm_weight, post_norm_bias, eps=1e-6, undo_index=None):
    """Output side before wo (ttt_layer.py:456,472,324,329-331): op_out bf16 [B,H,NC,CS,64] -> post_norm(transpose) in the
    caller's token order, bf16 [B,L,H*64].  undo_index: int32 [L] gather index of undo_interleave (None for one scene).
    Differentiable w.r.t. op_out, post_norm_weight, post_norm_bias."""
    return _OutputNorm.apply(op_out, post_norm_weight, post_norm_bias, eps, undo_index)
