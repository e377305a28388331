"""CPU (build container only for the first test): ``sampling.BatchedDenoiser`` reproduces the reference's
``DiscreteDenoiser.forward`` (ttt/models/cogvideo/utils.py:460-492) on the classifier-free-guidance pair with ONE network
call instead of one per batch element."""
import os
import sys

import pytest
import torch

REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "ttt")), reason="reference tree not present")
def test_batched_denoiser_matches_reference_loop():
    sys.path.insert(0, REF)
    try:
        from ttt.models.cogvideo import utils as RU
    except Exception as e:  # optional dependencies of the reference's utils (wandb, tqdm, ...) missing
        pytest.skip(f"reference utils not importable: {e}")
    finally:
        sys.path.remove(REF)
    from ttt_video_dit_b200.sampling import BatchedDenoiser
    torch.manual_seed(0)
    calls = []

    def network(x, crossattn, c_noise):  # per-element results must not depend on the batching
        calls.append(x.shape[0])
        return torch.tanh(x) * (1 + 0.01 * c_noise.float().reshape(-1, 1, 1, 1, 1)) + crossattn.mean(dim=(1, 2)).reshape(-1, 1, 1, 1, 1)
    ref = RU.DiscreteDenoiser.__new__(RU.DiscreteDenoiser)
    torch.nn.Module.__init__(ref)
    ref.scaling = RU.VideoScaling()
    ref.sigmas = RU.ZeroSNRDDPMDiscretization()(1000, do_append_zero=False, device="cpu", flip=True)
    ref.quantize_c_noise, ref.network, ref.dtype = True, network, torch.float32
    mine = BatchedDenoiser(network, ref.sigmas, dtype=torch.float32)
    x = torch.randn(2, 3, 4, 5, 6)
    sigma = ref.sigmas[[400, 400]].clone()
    cond = {"crossattn": torch.randn(2, 7, 8)}
    idx = torch.tensor([400.0, 400.0])
    a = ref(x, sigma, cond, idx=ref.sigmas[[400, 400]])
    n_ref = len(calls)
    b = mine(x, sigma, cond, idx=ref.sigmas[[400, 400]])
    assert calls[:n_ref] == [1, 1] and calls[n_ref:] == [2]
    assert torch.allclose(a, b, rtol=0, atol=1e-6)


def test_sampling_ops_need_the_cuda_library_and_device():
    from ttt_video_dit_b200 import sampling
    from ttt_video_dit_b200.transformer_layer import LayerMeta
    meta = LayerMeta(num_heads=2, text_length=16, num_chunks=1, num_frames=13, latent_height=4, latent_width=4)
    with pytest.raises(RuntimeError):
        sampling.dit_stack_forward(torch.zeros(1, 64, 128, dtype=torch.bfloat16), torch.zeros(1, 8), [{}], meta)
