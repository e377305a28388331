"""Multi-scene token order as gather indices (host logic; no kernel here).

The reference reorders tokens around the scan with two tensor-slicing routines, ``TTTBase.interleave`` (ttt/models/ssm/
ttt_layer.py:157-189: [all text | all video] -> [text_0 video_0 | text_1 video_1 | ...]) and ``TTTBase.undo_interleave``
(:191-217, the inverse on the layer output).  The kernels of this package fuse both permutations into passes that already
touch every token (csrc/process_input.cu, csrc/output_norm.cu) and therefore take them as int32 gather indices:

    interleaved[:, l] = x[:, interleave_index[l]]          undone[:, m] = y[:, undo_interleave_index[m]]

Scene layout (SequenceMetadata fields of the reference): ``text_length`` tokens of text per scene, ``num_chunks`` scenes,
scene 0 spans ``init_offset`` tokens (text + one extra latent frame of video), every later scene ``base_offset`` tokens.
"""
from typing import Optional, Tuple

import torch


def _video_parts(n_video: int, first: int, num_chunks: int):
    """Lengths of the per-scene video parts: `first` tokens for scene 0, the rest cut like torch.chunk(num_chunks - 1)."""
    rest = n_video - first
    if first < 0 or rest < 0:
        raise ValueError("interleave: init_offset does not fit the sequence")
    parts = [first]
    if num_chunks > 1:
        size = -(-rest // (num_chunks - 1)) if rest else 0  # torch.chunk: ceil-sized pieces, a shorter (or missing) tail
        while rest > 0:
            parts.append(min(size, rest))
            rest -= parts[-1]
        if len(parts) != num_chunks:
            raise ValueError("interleave: video tokens do not split into num_chunks - 1 scene parts")
    elif rest:
        raise ValueError("interleave: one scene must hold all video tokens")
    return parts


def interleave_index(L: int, text_length: int, num_chunks: int, init_offset: int, device=None) -> torch.Tensor:
    """Gather index of ``TTTBase.interleave`` over the flattened token axis, int32 [L]."""
    seq_text = text_length * num_chunks
    if seq_text > L:
        raise ValueError("interleave: more text tokens than tokens")
    parts = _video_parts(L - seq_text, init_offset - text_length, num_chunks)
    idx = torch.empty(L, dtype=torch.int32)
    out, vid = 0, seq_text
    for i, n in enumerate(parts):
        idx[out:out + text_length] = torch.arange(i * text_length, (i + 1) * text_length, dtype=torch.int32)
        out += text_length
        idx[out:out + n] = torch.arange(vid, vid + n, dtype=torch.int32)
        out += n
        vid += n
    return idx if device is None else idx.to(device)


def undo_interleave_index(L: int, text_length: int, num_chunks: int, init_offset: int, base_offset: int,
                          device=None) -> torch.Tensor:
    """Gather index of ``TTTBase.undo_interleave``: scene-ordered tokens back to [all text | all video], int32 [L]."""
    starts = [0] + [init_offset + i * base_offset for i in range(num_chunks - 1)]
    ends = [init_offset] + [init_offset + (i + 1) * base_offset for i in range(num_chunks - 1)]
    if ends[-1] != L or any(e - s < text_length for s, e in zip(starts, ends)):
        raise ValueError("undo_interleave: scene offsets do not tile the sequence")
    text = [torch.arange(s, s + text_length, dtype=torch.int32) for s in starts]
    video = [torch.arange(s + text_length, e, dtype=torch.int32) for s, e in zip(starts, ends)]
    idx = torch.cat(text + video)
    return idx if device is None else idx.to(device)


def indices_from_metadata(L: int, seq_metadata, device=None) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """(interleave_index, undo_interleave_index) for a reference ``SequenceMetadata`` (ttt/models/cogvideo/utils.py:220-238), or
    (None, None) when the sequence is a single scene (``is_multiscene`` false: the reference skips both routines,
    ttt_layer.py:290-294, 329-331)."""
    if not getattr(seq_metadata, "is_multiscene", False):
        return None, None
    tl, n, io, bo = seq_metadata.text_length, seq_metadata.num_chunks, seq_metadata.init_offset, seq_metadata.base_offset
    if io is None or bo is None:
        raise ValueError("interleave: init_offset and base_offset must be set for a multi-scene sequence")
    return interleave_index(L, tl, n, io, device), undo_interleave_index(L, tl, n, io, bo, device)
