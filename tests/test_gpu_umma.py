"""GPU: tcgen05 descriptor conventions (csrc/umma_selftest.cu) against torch fp32 matmul of the same bf16 operands."""
import pytest
import torch

from ttt_video_dit_b200 import _lib

pytestmark = pytest.mark.gpu


def run_umma(mode, N, K, seed=0):
    g = torch.Generator().manual_seed(seed)
    f16 = mode >= 9
    A = torch.randn(128, K, generator=g).to(torch.float16 if mode in (4, 5) or f16 else torch.bfloat16).cuda()
    if mode == 6:
        A[64:] = A[:64]  # the kernel stages only rows 0-63; LBO = 0 must replicate them
    Bm = torch.randn(K, N, generator=g).to(torch.float16 if f16 else torch.bfloat16).cuda()
    D = torch.zeros(128, N, device="cuda")
    b_arg = Bm.t().contiguous() if mode == 3 else Bm
    code = _lib.debug_lib().ttt_b200_debug_umma(mode, _lib.ptr(A), _lib.ptr(b_arg), _lib.ptr(D), N, K, _lib.current_stream(A))
    assert code == 0, (code, _lib.debug_lib().ttt_b200_debug_last_error())
    torch.cuda.synchronize()
    ref = A.float() @ Bm.float()
    return float((D - ref).abs().max() / ref.abs().max())


@pytest.mark.parametrize("mode,N,K", [(0, 64, 64), (0, 128, 64), (0, 64, 256), (1, 64, 64), (1, 64, 256),
                                      (2, 64, 64), (2, 64, 128), (3, 64, 64), (3, 128, 64), (6, 64, 256),
                                      (7, 64, 64), (7, 128, 64), (7, 64, 128), (8, 64, 64), (8, 64, 128),
                                      (9, 64, 64), (9, 128, 64), (10, 64, 256), (11, 64, 64)])  # 9-11: both operands fp16  # 7/8: A operand from TMEM  # modes 4/5 (A=f16 with B=bf16) trap with 'illegal instruction' on sm_100a: mixed formats are not allowed
def test_umma_modes(mode, N, K):
    assert run_umma(mode, N, K) < 1e-5
