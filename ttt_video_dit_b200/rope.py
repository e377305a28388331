"""3-D rotary position tables of the DiT block (host logic, computed once per sequence shape).

Two consumers with two conventions, both from the reference:
  * the local attention rotates interleaved feature pairs with per-position sin / cos tables of width head_dim in which
    each angle appears twice (``Rotary3DPositionEmbedding``, ttt/models/cogvideo/utils.py:388-437: time / height / width
    bands of head_dim/4, 3*head_dim/8, 3*head_dim/8 features, positions local to the segment);
  * the TTT layer rotates the same pairs from complex ``freqs_cis`` (``precompute_freqs_cis_3d``, ttt/models/ssm/utils.py:9-53)
    -- taken by csrc/process_input.cu as (cos, sin) tables of width head_dim/2 over the GLOBAL video positions.
"""
import torch


def _band_angles(num_frames, height, width, head_dim, theta):
    bands = (head_dim // 4, head_dim // 8 * 3, head_dim // 8 * 3)  # time, height, width
    grids = []
    for n, dim in zip((num_frames, height, width), bands):
        inv = 1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
        grids.append(torch.arange(n, dtype=torch.float32)[:, None] * inv[None])
    return grids


def _broadcast(gt, gh, gw):
    T, H, W = gt.shape[0], gh.shape[0], gw.shape[0]
    return torch.cat([gt[:, None, None, :].expand(T, H, W, -1), gh[None, :, None, :].expand(T, H, W, -1),
                      gw[None, None, :, :].expand(T, H, W, -1)], dim=-1).reshape(T * H * W, -1)


def attention_tables(height, width, num_frames, head_dim, theta=10000.0, device=None):
    """(sin, cos) [(t h w), head_dim] for attention.local_attention."""
    gt, gh, gw = (g.repeat_interleave(2, dim=-1) for g in _band_angles(num_frames, height, width, head_dim, theta))
    ang = _broadcast(gt, gh, gw)
    sin, cos = ang.sin(), ang.cos()
    return (sin, cos) if device is None else (sin.to(device), cos.to(device))


def ttt_tables(height, width, num_frames, head_dim, theta=10000.0, device=None):
    """(cos, sin) [(t h w), head_dim / 2] for process_input (TTT q / k rotation)."""
    ang = _broadcast(*_band_angles(num_frames, height, width, head_dim, theta))
    cos, sin = ang.cos(), ang.sin()
    return (cos, sin) if device is None else (cos.to(device), sin.to(device))
