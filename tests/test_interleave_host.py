"""CPU: the product-side gather indices (ttt_video_dit_b200.interleave) reproduce the reference's tensor-slicing routines
TTTBase.interleave / undo_interleave (ttt/models/ssm/ttt_layer.py:157-217), replayed here on a tensor of token ids, and
agree with the oracle's restatement; scene offsets as get_interleave_offsets (ttt/models/cogvideo/utils.py:16-26) makes them."""
import types

import pytest
import torch
from hypothesis import given, settings, strategies as st

from oracle import ttt_oracle as O
from ttt_video_dit_b200 import interleave as I


def offsets(num_frames, num_chunks, tokens_per_frame, text_length):
    fpc = num_frames // num_chunks
    return fpc * tokens_per_frame + text_length, (fpc + num_frames % fpc) * tokens_per_frame + text_length  # base, init


def replay_interleave(ids, text_length, num_chunks, init_offset):
    """The reference's slicing (ttt_layer.py:165-188) applied to token ids [L]."""
    seq_text = text_length * num_chunks
    text = torch.chunk(ids[:seq_text], num_chunks)
    video = ids[seq_text:]
    v0 = init_offset - text_length
    video = (video[:v0],) + torch.chunk(video[v0:], num_chunks - 1)
    return torch.cat([torch.cat((text[i], video[i])) for i in range(num_chunks)])


def replay_undo(ids, text_length, num_chunks, init_offset, base_offset):
    """ttt_layer.py:205-217 applied to token ids [L]."""
    text, vid = [], []
    for i in range(num_chunks):
        s = 0 if i == 0 else init_offset + (i - 1) * base_offset
        e = init_offset if i == 0 else init_offset + i * base_offset
        text.append(ids[s:e][:text_length])
        vid.append(ids[s:e][text_length:])
    return torch.cat(text + vid)


CASES = [(13, 1, 8, 6), (37, 3, 8, 6), (37, 3, 1350, 502), (73, 6, 12, 7), (253, 21, 4, 3), (25, 2, 16, 16)]


@pytest.mark.parametrize("frames,chunks,tpf,tl", CASES)
def test_indices_match_reference_slicing_and_oracle(frames, chunks, tpf, tl):
    if chunks == 1:
        md = types.SimpleNamespace(is_multiscene=False)
        assert I.indices_from_metadata(frames * tpf + tl, md) == (None, None)
        return
    base, init = offsets(frames, chunks, tpf, tl)
    L = chunks * tl + frames * tpf
    ids = torch.arange(L)
    idx = I.interleave_index(L, tl, chunks, init)
    und = I.undo_interleave_index(L, tl, chunks, init, base)
    assert idx.dtype == torch.int32 and und.dtype == torch.int32 and idx.shape == und.shape == (L,)
    assert torch.equal(ids[idx.long()], replay_interleave(ids, tl, chunks, init))
    assert torch.equal(ids[und.long()], replay_undo(ids, tl, chunks, init, base))
    assert torch.equal(idx.long(), O.interleave_index(L, tl, chunks, init))
    assert torch.equal(und.long(), O.undo_interleave_index(L, tl, chunks, init, base))
    assert torch.equal(idx.long()[und.long()], ids)  # undo o interleave = identity for offsets the reference produces
    md = types.SimpleNamespace(is_multiscene=True, text_length=tl, num_chunks=chunks, init_offset=init, base_offset=base)
    a, b = I.indices_from_metadata(L, md)
    assert torch.equal(a, idx) and torch.equal(b, und)


@settings(max_examples=60, deadline=None)
@given(chunks=st.integers(2, 9), fpc=st.integers(1, 6), extra=st.integers(0, 5), tpf=st.integers(1, 9), tl=st.integers(0, 11))
def test_indices_are_inverse_permutations(chunks, fpc, extra, tpf, tl):
    frames = fpc * chunks + (extra % fpc)
    base, init = offsets(frames, chunks, tpf, tl)
    L = chunks * tl + frames * tpf
    ids = torch.arange(L)
    idx = I.interleave_index(L, tl, chunks, init).long()
    und = I.undo_interleave_index(L, tl, chunks, init, base).long()
    assert torch.equal(torch.sort(idx).values, ids) and torch.equal(torch.sort(und).values, ids)
    assert torch.equal(ids[idx], replay_interleave(ids, tl, chunks, init))
    assert torch.equal(idx[und], ids)


def test_inconsistent_layouts_raise():
    with pytest.raises(ValueError):
        I.interleave_index(100, 60, 2, 70)            # more text than tokens
    with pytest.raises(ValueError):
        I.interleave_index(100, 10, 2, 5)             # init_offset shorter than the text of scene 0
    with pytest.raises(ValueError):
        I.undo_interleave_index(100, 10, 3, 40, 25)   # 40 + 2*25 != 100
    with pytest.raises(ValueError):
        I.indices_from_metadata(100, types.SimpleNamespace(is_multiscene=True, text_length=10, num_chunks=2,
                                                           init_offset=None, base_offset=None))
