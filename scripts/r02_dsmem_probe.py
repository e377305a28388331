"""GPU: DSMEM transfer cost between the two CTAs of a cluster (see csrc/dsmem_probe.cu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ttt_video_dit_b200 import _lib
out = torch.zeros(4, device="cuda")
L = _lib.debug_lib()
for mode, name in ((0, "bulk copy"), (2, "4 bulk copies"), (1, "st.shared::cluster x256 threads")):
    for nbytes in (64, 1024, 4096, 8192, 16384, 32768, 65536):
        rc = L.ttt_b200_debug_dsmem(mode, nbytes, 200, _lib.ptr(out), None)
        torch.cuda.synchronize()
        assert rc == 0, (rc, L.ttt_b200_debug_last_error())
        cyc = float(out[0])
        print(f"{name:34s} {nbytes:6d} B: {cyc:8.0f} cycles one way  ({nbytes / cyc:6.1f} B/cycle incl. latency)", flush=True)
