"""CPU: the product ops have no eager / CPU path.  Called with host tensors every public entry raises instead of computing,
and with the shared library hidden the loader raises instead of degrading -- what makes a GPU test that passes mean that
the CUDA kernels ran."""
import os
import subprocess
import sys

import pytest
import torch

from oracle import ttt_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
bf = lambda t: t.to(torch.bfloat16)


def test_every_public_op_refuses_host_tensors():
    from ttt_video_dit_b200 import attention, linear_triton, mlp_tk, process_input, seq_block, transformer_layer
    d = O.make_inputs(1, 2, 2, seed=0)
    prm = [d[k] for k in ("ln_w", "ln_b", "W1", "b1", "W2", "b2")]
    dl = O.make_inputs(1, 2, 2, CS=16, seed=0, base_lr=1.0, linear=True)
    q = bf(torch.randn(1, 128, 2, 64))
    x = bf(torch.randn(1, 128, 128))
    gate = torch.full((128,), 0.1)
    calls = {
        "TkMLP": lambda: mlp_tk.TkMLP.apply(*prm, bf(d["XQ"]), bf(d["XV"]), bf(d["XK"]), bf(d["eta"]), 1),
        "ttt_mlp_op": lambda: mlp_tk.ttt_mlp_op(*prm, bf(d["XQ"]), bf(d["XV"]), bf(d["XK"]), bf(d["eta"][:, :, :, -1:, :]), 1),
        "TritonLinear": lambda: linear_triton.TritonLinear.apply(dl["ln_w"], dl["ln_b"], dl["W1"], dl["b1"], bf(dl["XQ"]),
                                                                 bf(dl["XV"]), bf(dl["XK"]), bf(dl["eta"]), 1),
        "sdpa_bthd": lambda: attention.sdpa_bthd(q, q, q),
        "prepare": lambda: process_input.prepare(x, x, x, torch.randn(1, 128, 2), torch.randn(128, 32), torch.randn(128, 32),
                                                 torch.ones(2, 64), torch.zeros(2, 64), 0, 64, 0.1),
        "output_norm": lambda: process_input.output_norm(bf(torch.randn(1, 2, 2, 64, 64)), torch.ones(128), torch.zeros(128)),
        "ssm_forward": lambda: seq_block.ssm_forward(x, lambda t: t, 0, 1, False, gate, gate, gate, gate),
        "LnAffine": lambda: transformer_layer.LnAffine.apply(x, torch.ones(1, 2, 128), torch.zeros(1, 2, 128), 0, 1e-6),
        "QKNormRope": lambda: attention.QKNormRope.apply(q, q, torch.ones(64), torch.zeros(64), torch.ones(64), torch.zeros(64),
                                                         torch.zeros(128, 64), torch.ones(128, 64), 0, 1e-6),
        "GateAdd": lambda: transformer_layer.GateAdd.apply(x, x, torch.ones(1, 2, 128), 0),
    }
    for name, fn in calls.items():
        with pytest.raises(RuntimeError, match="CUDA"):
            fn()


def test_missing_library_raises_at_load():
    code = ("import os, sys; sys.path.insert(0, %r); os.environ['TTT_B200_LIB'] = '/nonexistent/libttt_b200.so'\n"
            "from ttt_video_dit_b200 import _lib\n"
            "try:\n    _lib.lib()\nexcept Exception as e:\n    print('RAISED', type(e).__name__)\nelse:\n    print('LOADED')\n") % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "RAISED" in r.stdout, r.stdout + r.stderr[-1000:]
