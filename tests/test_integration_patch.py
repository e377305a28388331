"""CPU (build container only; skipped where /root/reference is absent): integration/ttt_video_dit_b200.patch applies to the
reference tree, binds the reference's TkMLP / TritonLinear names to this package's classes, and leaves the reference's own
CPU (eager, fp32/fp64) path bit-identical -- the B200 branches are taken for CUDA bf16 tensors only."""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
PATCH = os.path.join(ROOT, "integration", "ttt_video_dit_b200.patch")

SCRIPT = r'''
import sys, torch
torch.manual_seed(0)
from ttt.models.cogvideo.dit import SeqModelingBlock
from ttt.models.cogvideo.utils import SequenceMetadata
from ttt.models.configs import ModelConfig
import ttt.models.ssm.ttt_layer as TL
print("TKMLP", TL.TkMLP.__module__, "LINEAR", TL.TritonLinear.__module__)
E, NH, Hh, Ww, frames, T, chunks = 128, 2, 4, 4, 25, 8, 2
cfg = ModelConfig(model_dim=E, num_heads=NH, num_layers=1, ssm_layer="ttt_linear", mini_batch_size=16, ttt_base_lr=1.0,
                  latent_height=Hh, latent_width=Ww, compressed_num_frames=frames, adapter_method="sft")
blk = SeqModelingBlock(cfg).double()
class Stub(torch.nn.Module):
    def forward(self, x, seq_metadata):
        return torch.cumsum(x, dim=1) * 0.01 + torch.roll(x, 1, dims=-1) * 0.5
blk.ssm = Stub()
md = SequenceMetadata(text_length=T, seq_text_length=T * chunks, num_frames=frames, num_chunks=chunks, tokens_per_frame=Hh * Ww,
                      latent_height=Hh, latent_width=Ww, t_emb=torch.zeros(1, 8))
md.init_multiscene_offsets()
vid = torch.randn(2, frames * Hh * Ww, E, dtype=torch.float64)
txt = torch.randn(2, T * chunks, E, dtype=torch.float64)
with torch.no_grad():
    a = blk._attn_forward(vid, txt, md)
    s = blk._ssm_forward(torch.cat((txt, vid), dim=1), md)
print("SUM %.12e %.12e" % (a.double().sum().item(), s.double().sum().item()))
'''


def run(pythonpath):
    env = dict(os.environ, TORCHDYNAMO_DISABLE="1", PYTHONPATH=os.pathsep.join(pythonpath))
    r = subprocess.run([sys.executable, "-c", SCRIPT], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    return r.stdout


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "ttt")), reason="reference tree not present")
def test_patch_applies_binds_and_keeps_the_cpu_path(tmp_path):
    shutil.copytree(os.path.join(REF, "ttt"), tmp_path / "ttt")
    r = subprocess.run(["patch", "-p1", "--no-backup-if-mismatch", "-i", PATCH], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    patched = run([str(tmp_path), ROOT])
    stock = run([REF])
    assert "TKMLP ttt_video_dit_b200.mlp_tk LINEAR ttt_video_dit_b200.linear_triton" in patched
    assert "TKMLP ttt.models.ssm.mlp_tk LINEAR ttt.models.ssm.linear_triton" in stock
    assert [l for l in patched.splitlines() if l.startswith("SUM")] == [l for l in stock.splitlines() if l.startswith("SUM")]
