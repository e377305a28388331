"""Sampling-loop integration of the hot path (BASELINE config 5: 63-second video, 50-step forward-only).

The reference's sampler calls the DiT once per batch ELEMENT: ``DiscreteDenoiser.forward`` (ttt/models/cogvideo/utils.py:460-492)
loops ``for i in range(batch_size)`` over the classifier-free-guidance pair that ``DynamicCFG.prepare_inputs`` (:528-537) has
just concatenated, so one video costs 100 batch-1 DiT forwards (50 steps x {unconditional, conditional}).  On a B200 a
batch-1 TTT scan fills 48 of 148 SMs; the guidance pair as ONE batch of 2 runs 96 scan CTAs concurrently and halves the number
of launches.  This module is that loop, nothing else of the sampler (schedules, the DPM++ step and the VAE are out of scope):

  * ``BatchedDenoiser``   -- same arithmetic and call signature as ``DiscreteDenoiser.forward``, one network call per step;
  * ``dit_stack_forward`` -- forward-only stack of ``transformer_layer_forward`` layers (no autograd, one checkpoint group:
                             the eval configs set scan_checkpoint_group_size = 1e6, configs/eval/ttt-mlp/63s.toml:44-45);
  * ``GraphedCall``       -- CUDA-graph capture of a shape-static callable (the 50 steps replay the same launch sequence;
                             a forward of the 42-layer stack is ~2 000 kernel launches).
"""
from typing import Callable, Dict, Sequence

import torch

from . import _lib
from .transformer_layer import LayerMeta, transformer_layer_forward


def _append_dims(x, ndim):
    return x[(...,) + (None,) * (ndim - x.ndim)]


class BatchedDenoiser(torch.nn.Module):
    """``DiscreteDenoiser`` (cogvideo/utils.py:441-509) with the per-element loop of ``forward`` replaced by one batched call.

    network(scaled_input [B, ...], crossattn [B, ...], c_noise [B]) -> [B, ...];  sigmas: the discretisation table the
    reference builds in its constructor (``ZeroSNRDDPMDiscretization()(num_idx, flip=True)``), passed in by the caller."""

    def __init__(self, network: Callable, sigmas: torch.Tensor, dtype=torch.bfloat16, quantize_c_noise: bool = True):
        super().__init__()
        self.network, self.dtype, self.quantize_c_noise = network, dtype, quantize_c_noise
        self.register_buffer("sigmas", sigmas.clone(), persistent=False)

    def sigma_to_idx(self, sigma):  # utils.py:494-496
        dists = sigma - self.sigmas.to(sigma.device)[:, None]
        return dists.abs().argmin(dim=0).view(sigma.shape)

    def idx_to_sigma(self, idx):  # utils.py:498-499
        return self.sigmas.to(idx.device)[idx]

    def forward(self, input: torch.Tensor, sigma: torch.Tensor, cond: Dict, idx: torch.Tensor = None, **unused):
        sigma = self.idx_to_sigma(self.sigma_to_idx(sigma))              # possibly_quantize_sigma (:501-502)
        sigma_shape = sigma.shape
        sigma = _append_dims(sigma, input.ndim)
        c_skip, c_out = sigma, -((1 - sigma ** 2) ** 0.5)                # VideoScaling (:252-258): c_in = 1, c_noise = idx
        c_noise = idx.clone().reshape(sigma_shape)
        if self.quantize_c_noise:
            c_noise = self.sigma_to_idx(c_noise)
        scaled = input.to(dtype=self.dtype)
        out = self.network(scaled, cond["crossattn"], c_noise)           # ONE call for the whole (guidance) batch
        return out * c_out + input * c_skip


@torch.no_grad()
def dit_stack_forward(emb: torch.Tensor, t_emb: torch.Tensor, layers: Sequence[Dict[str, torch.Tensor]], meta: LayerMeta):
    """emb bf16 [B, L, E] (text tokens first) through ``len(layers)`` TransformerLayers (state_dicts with the reference's
    names), forward only.  The scans keep a single checkpoint (nothing is saved for a backward)."""
    _lib.lib()
    nc = emb.shape[1] // meta.mini_batch_size
    saved = meta.scan_checkpoint_group_size
    meta.scan_checkpoint_group_size = max(nc, 1)
    try:
        for P in layers:
            emb = transformer_layer_forward(emb, t_emb, P, meta)
    finally:
        meta.scan_checkpoint_group_size = saved
    return emb


class GraphedCall:
    """Capture ``fn(*static_inputs)`` once into a CUDA graph and replay it: ``out = graphed(*new_inputs)`` copies the new
    values into the captured input buffers, launches the graph and returns the captured output buffer (valid until the
    next call).  ``fn`` must be shape-static, allocate only through torch's caching allocator and launch on the current
    stream -- true of every op in this package (no device synchronisation, no host-side data-dependent control flow)."""

    def __init__(self, fn: Callable, *example_inputs: torch.Tensor, warmup: int = 2):
        self.static_in = [t.clone() for t in example_inputs]
        side = torch.cuda.Stream(self.static_in[0].device)
        side.wait_stream(torch.cuda.current_stream(self.static_in[0].device))
        with torch.cuda.stream(side):  # lazily-initialised state (function attributes, cuBLAS handles, streams) before the capture
            for _ in range(warmup):
                fn(*self.static_in)
        torch.cuda.current_stream(self.static_in[0].device).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.static_out = fn(*self.static_in)

    def __call__(self, *inputs: torch.Tensor):
        for dst, src in zip(self.static_in, inputs):
            dst.copy_(src)
        self.graph.replay()
        return self.static_out
