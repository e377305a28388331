"""GPU parity of the adaLN shell kernels (csrc/adaln.cu: ln_affine = LayerNorm + modulate, gate_add = gated residual; forward
and backward) vs torch autograd of the reference formulas (ttt/models/cogvideo/dit.py:344-350, utils.py:70-75) in fp32."""
import pytest
import torch

from oracle import ttt_oracle as O
from ttt_video_dit_b200 import transformer_layer as TL

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,L,E,Lt", [(1, 70, 128, 6), (2, 333, 3072, 100), (1, 64, 3072, 0), (2, 33, 256, 33)])
def test_ln_affine_and_gate_add(B, L, E, Lt):
    g = torch.Generator().manual_seed(L + E)
    rn = lambda *s: torch.randn(*s, generator=g)
    x, y, go = rn(B, L, E).bfloat16(), rn(B, L, E).bfloat16(), rn(B, L, E).bfloat16()
    gamma, beta = 1 + 0.2 * rn(E), 0.2 * rn(E)
    mods = [0.3 * rn(B, E) for _ in range(6)]  # shift, scale, gate (video), then text
    eps = 1e-6

    def ref():
        xl, yl = x.float().requires_grad_(True), y.float().requires_grad_(True)
        gm, bt = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
        m = [t.clone().requires_grad_(True) for t in mods]
        ln = torch.nn.functional.layer_norm(xl, (E,), gm, bt, eps)
        mod = lambda t, sh, sc: t * (1 + sc[:, None]) + sh[:, None]            # modulate (cogvideo/utils.py:70-75)
        h = torch.cat((mod(ln[:, :Lt], m[3], m[4]), mod(ln[:, Lt:], m[0], m[1])), dim=1)
        gate = torch.cat((m[5][:, None].expand(B, Lt, E), m[2][:, None].expand(B, L - Lt, E)), dim=1)
        out = h + gate * yl                                                        # emb + gate * block_out (dit.py:349-350)
        out.backward(go.float())
        return out.detach(), [xl.grad, yl.grad, gm.grad, bt.grad] + [t.grad for t in m]

    xd, yd = x.cuda().requires_grad_(True), y.cuda().requires_grad_(True)
    gm, bt = gamma.cuda().requires_grad_(True), beta.cuda().requires_grad_(True)
    m = [t.cuda().requires_grad_(True) for t in mods]
    A, C = TL._affine(gm, bt, m[0], m[1], m[3], m[4])
    h = TL.LnAffine.apply(xd, A, C, Lt, eps)
    out = TL.GateAdd.apply(h, yd, torch.stack((m[5], m[2]), dim=1), Lt)
    out.backward(go.cuda())
    torch.cuda.synchronize()
    r_out, r_g = ref()
    assert O.rel_err(out.float().cpu(), r_out) < 1e-2
    got = [xd.grad, yd.grad, gm.grad, bt.grad] + [t.grad for t in m]
    names = ["dx", "dy", "dgamma", "dbeta", "dshift", "dscale", "dgate", "dtshift", "dtscale", "dtgate"]
    for n, a, b in zip(names, got, r_g):
        if b.abs().max() == 0:      # no text (or no video) rows: the gradient of that segment's vectors is exactly zero
            assert a.abs().max() == 0, n
            continue
        assert O.rel_err(a.float().cpu(), b) < 2e-2, (n, O.rel_err(a.float().cpu(), b))
