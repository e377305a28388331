"""Diagnostic runner for the GPU box: each probe in its own process (a trapped kernel poisons the CUDA context)."""
import subprocess
import sys
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PROBES = {
    "umma": """
import torch, sys
sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests')
from test_gpu_umma import run_umma
for mode, N, K in [(0,64,64),(7,64,64),(7,128,64),(7,64,128),(8,64,64),(8,64,128),(6,64,256)]:
    try:
        print('umma mode', mode, 'N', N, 'K', K, 'relmax', run_umma(mode, N, K), flush=True)
    except Exception as e:
        print('umma mode', mode, N, K, 'EXC', repr(e)[:300], flush=True); break
""" % (ROOT, ROOT),
    "fwd": """
import torch, sys
sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests')
from oracle import ttt_oracle as O
from test_gpu_mlp_forward import run_forward, oracle_forward
for (B,H,NC,G) in [(1,1,1,1),(1,1,2,1),(1,2,4,2),(2,3,7,3),(1,4,33,16)]:
    d = O.make_inputs(B,H,NC,seed=10+NC)
    qkve, out, ck, last = run_forward(d, G, want_last=True)
    ref, rck, rlast = oracle_forward(qkve, d, G)
    per = [O.rel_err(out[:,:,n].float().cpu(), ref[:,:,n]) for n in range(min(NC,4))]
    print('fwd', (B,H,NC,G), 'out', O.rel_err(out.float().cpu(), ref), 'per-step', per,
          'ck', [O.rel_err(a.cpu(), b) for a,b in zip(ck, rck)], 'last', [O.rel_err(a.cpu(), b) for a,b in zip(last, rlast)], flush=True)
""" % (ROOT, ROOT),
    "timing": """
import os
os.environ['TTT_B200_LIB'] = os.environ.get('TTT_B200_TIMING_LIB', %r + '/ttt_video_dit_b200/lib/libttt_b200_dbg.so')
import torch, sys
sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests')
from oracle import ttt_oracle as O
from ttt_video_dit_b200 import _lib, mlp_tk
buf = torch.zeros(16384, dtype=torch.int32, device='cuda')
print('timing build:', _lib.set_timing_buffer(buf))
B,H,NC,G = 1,int(os.environ.get('TTT_TIMING_H','48')),int(os.environ.get('TTT_TIMING_NC','64')),16
d = O.make_inputs(B,H,NC,seed=1)
bf = lambda t: t.to(torch.bfloat16).cuda()
prm = [d[k].cuda().requires_grad_(True) for k in ('ln_w','ln_b','W1','b1','W2','b2')]
q,v,k = [bf(d[n]).requires_grad_(True) for n in ('XQ','XV','XK')]
e = bf(d['eta'])[:,:,:,-1,:].clone().requires_grad_(True)
for rep in range(2):
    out = mlp_tk.ttt_mlp_op(*prm, q, v, k, e, G)
    torch.cuda.synchronize()
    t = buf.cpu().tolist()
    names = ['top','tma_wait','P1 mma','P2 gelu','P3 mma','P4 LN','P5 mma','P6 gradZ1','P7 mma','P8 remat']
    if rep == 1:
        for ob in (0,1):
            tot = sum(t[ob*64:ob*64+10])
            print('FWD observer', ob, 'cycles/step total', tot/NC, {n: round(t[ob*64+i]/NC) for i,n in enumerate(names)}, flush=True)
    buf.zero_()
    out.backward(bf(d['dOut']))
    torch.cuda.synchronize()
    t = buf.cpu().tolist()
    bn = ['top','apply+A0 cw','A1 mma','A2 gelu3','A3 mma','A4 tok','A56 mma','A56 ew','A7 mma','A8 tok','A9 mma','A10 ew','A11 mma','A12 tok','PROLOGUE(total)','looptail','EPILOGUE(total)']
    if rep == 1:
        for ob in (0,1):
            tot = sum(t[ob*64:ob*64+14])
            div = NC if os.environ.get('TTT_B200_PERSISTENT', '1') != '0' else G  # persistent: the slots accumulate over the whole scan
            print('   prologue cycles', t[ob*64+14], 'epilogue cycles', t[ob*64+16], flush=True)
            print('BWD observer', ob, 'cycles/step total', tot/div, {n: round(t[ob*64+i]/div) for i,n in enumerate(bn[:14])}, flush=True)
        print('per-block K-kernel cycles (sorted):', sorted(t[128:128+48]), flush=True)
        print('smid of block b:', t[192:192+48], flush=True)
        ngr = (NC + G - 1) // G
        base = None
        for gi in range(min(ngr, 27) - 1, -1, -1):
            st = [x & 0xffffffff for x in t[512+gi*128:512+gi*128+min(H,48)]]; en = [x & 0xffffffff for x in t[512+gi*128+64:512+gi*128+64+min(H,48)]]
            if base is None: base = min(st)
            ss = sorted((x-base)/1e3 for x in st); ee = sorted((x-base)/1e3 for x in en)
            print('group %%2d: first start %%8.1f last start %%8.1f | first end %%8.1f last end %%8.1f  (us)' %% (gi, ss[0], ss[-1], ee[0], ee[-1]), flush=True)
        sts = [x & 0xffffffff for x in t[4096:4096+NC]]
        dur = [round(((sts[k-1]-sts[k]) & 0xffffffff)/1e3,1) for k in range(NC-1,0,-1)]
        print('step durations us (t = NC-1 .. 1; launch boundaries show as long steps):', dur, flush=True)
        print('sum of step durations us:', round(sum(dur),1), ' median', sorted(dur)[len(dur)//2], flush=True)
    buf.zero_()
""" % (ROOT, ROOT, ROOT),
    "timeline": """
import os
if os.environ.get('TTT_TIMELINE_DBG'): os.environ['TTT_B200_LIB'] = %r + '/ttt_video_dit_b200/lib/libttt_b200_dbg.so'
import torch, sys, json
sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests')
from oracle import ttt_oracle as O
from ttt_video_dit_b200 import mlp_tk
from torch.profiler import profile, ProfilerActivity
B,H,NC,G = 1,48,282,16
g = torch.Generator().manual_seed(0)
rn = lambda *s: torch.randn(*s, generator=g)
bf = lambda t: t.to(torch.bfloat16).cuda()
q = bf(torch.nn.functional.normalize(rn(B,H,NC,64,64), dim=-1)).requires_grad_(True)
k = bf(torch.nn.functional.normalize(rn(B,H,NC,64,64), dim=-1)).requires_grad_(True)
v = bf(rn(B,H,NC,64,64)).requires_grad_(True)
e = bf((0.1/64)*torch.sigmoid(rn(B,H,NC,64))/64).requires_grad_(True)
go = bf(rn(B,H,NC,64,64))
prm = [t.cuda().requires_grad_(True) for t in (1+0.1*rn(H,64), 0.1*rn(H,64), (0.02*rn(H,64,256)).unsqueeze(0).repeat(B,1,1,1), torch.zeros(B,H,1,256), (0.02*rn(H,256,64)).unsqueeze(0).repeat(B,1,1,1), torch.zeros(B,H,1,64))]
for _ in range(2):
    out = mlp_tk.ttt_mlp_op(*prm, q, v, k, e, G); out.backward(go)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    out = mlp_tk.ttt_mlp_op(*prm, q, v, k, e, G); out.backward(go)
    torch.cuda.synchronize()
evs = [ev for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA and 'ttt' in ev.name]
evs.sort(key=lambda ev: ev.time_range.start)
t0 = evs[0].time_range.start
for ev in evs[:70]:
    nm = 'FWD ' if 'fwd_kernelILb0' in ev.name else ('TRAJ' if 'fwd_kernelILb1' in ev.name else ('Q   ' if 'bwd_q' in ev.name else 'K   '))
    print(nm, 'start %%8.1f us  dur %%7.1f us' %% ((ev.time_range.start - t0), ev.time_range.end - ev.time_range.start), flush=True)
print('total span us', evs[-1].time_range.end - t0)
""" % (ROOT, ROOT, ROOT),
    "lin_bwd": """
import torch, sys
sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests')
from oracle import ttt_oracle as O
import test_gpu_linear_backward as T
for (B,H,NC,G) in [(1,1,1,1),(1,1,2,1),(1,2,3,2),(2,2,7,3),(1,3,20,16),(1,2,300,16)]:
    d = O.make_inputs(B,H,NC,CS=16,seed=70+NC,base_lr=1.0,linear=True)
    try:
        print((B,H,NC,G), {k: float('%%.2e' %% v) for k,v in T.errors(d,G).items()}, flush=True)
    except Exception as ex:
        print((B,H,NC,G), 'EXC', repr(ex)[:300], flush=True); break
import time
B,H,NC,G = 1,48,1128,16
d = O.make_inputs(B,H,NC,CS=16,seed=1,base_lr=1.0,linear=True)
dev='cuda'; bf = lambda t: t.to(torch.bfloat16).to(dev)
prm = [d[k].to(dev).requires_grad_(True) for k in ('ln_w','ln_b','W1','b1')]
q,v,k,e = [bf(d[n]).requires_grad_(True) for n in ('XQ','XV','XK','eta')]
go = bf(d['dOut'])
from ttt_video_dit_b200 import linear_triton
for rep in range(3):
    torch.cuda.synchronize(); t0=time.time()
    out = linear_triton.TritonLinear.apply(*prm, q, v, k, e, G)
    torch.cuda.synchronize(); t1=time.time()
    out.backward(go)
    torch.cuda.synchronize(); t2=time.time()
    print('linear NC=1128 B=1: fwd %%.3f ms  bwd %%.3f ms' %% ((t1-t0)*1e3, (t2-t1)*1e3), flush=True)
""" % (ROOT, ROOT),
    "lin_steps": """
import torch, sys
sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests')
from oracle import ttt_oracle as O
import test_gpu_linear_backward as T
for NC in (4, 6, 9):
    B,H,G = 1,1,1000
    d = O.make_inputs(B,H,NC,CS=16,seed=70+NC,base_lr=1.0,linear=True)
    out, g = T.run_fwd_bwd(d, G)
    r = lambda t: t.to(torch.bfloat16).float()
    le = r(d['eta'])[:,:,:,-1,:].unsqueeze(-1)
    ref = O.ttt_linear_primal_backward(r(d['XQ']), r(d['XK']), r(d['XV']), le, d['ln_w'], d['ln_b'], d['W1'], d['b1'], r(d['dOut']))
    for n, a, b in (('dXV', g[5], ref['dXV']), ('dXK', g[6], ref['dXK']), ('dXQ-dO', g[4].float().cpu() - r(d['dOut']), ref['dXQ'] - r(d['dOut'])), ('dEta', g[7][:,:,:,-1,:], ref['dlast_eta'].squeeze(-1))):
        a = a.float().cpu()
        print('NC', NC, n, 'per-step rel err', [float('%%.2e' %% O.rel_err(a[:,:,t], b[:,:,t])) for t in range(NC)], flush=True)
    print('NC', NC, 'dW1', O.rel_err(g[2].float().cpu(), ref['dW1']), 'db1', O.rel_err(g[3].float().cpu(), ref['db1']), flush=True)
""" % (ROOT, ROOT),
    "lin_timing": """
import os
os.environ['TTT_B200_LIB'] = %r + '/ttt_video_dit_b200/lib/libttt_b200_dbg.so'
import torch, sys
sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests')
from oracle import ttt_oracle as O
from ttt_video_dit_b200 import _lib, linear_triton
buf = torch.zeros(8192, dtype=torch.int32, device='cuda')
print('timing build:', _lib.set_timing_buffer(buf))
B,H,NC,G = 1,48,256,16
d = O.make_inputs(B,H,NC,CS=16,seed=1,base_lr=1.0,linear=True)
dev='cuda'; bf = lambda t: t.to(torch.bfloat16).to(dev)
prm = [d[k].to(dev).requires_grad_(True) for k in ('ln_w','ln_b','W1','b1')]
q,v,k,e = [bf(d[n]).requires_grad_(True) for n in ('XQ','XV','XK','eta')]
go = bf(d['dOut'])
for rep in range(2):
    out = linear_triton.TritonLinear.apply(*prm, q, v, k, e, G)
    buf.zero_()
    out.backward(go)
    torch.cuda.synchronize()
    t = buf.cpu().tolist()
    kn = ['waits K/img/Z','TP1','wait upd','cvt','issue C','D epilogue','wait dG','TP2','issue F']
    qn = ['waits Q/img','Zq mma','TPq','slot+tile','dQ mma','dQ store']
    if rep == 1:
        print('K group cycles/step total', sum(t[0:9])/NC, {n: round(t[i]/NC) for i,n in enumerate(kn)}, flush=True)
        print('Q group cycles/step total', sum(t[64:70])/NC, {n: round(t[64+i]/NC) for i,n in enumerate(qn)}, flush=True)
""" % (ROOT, ROOT, ROOT),
    "spin": """
import os
os.environ['TTT_B200_LIB'] = %r + '/ttt_video_dit_b200/lib/libttt_b200_dbg.so'
import torch, sys
sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests')
from oracle import ttt_oracle as O
from ttt_video_dit_b200 import _lib, mlp_tk
buf = torch.zeros(8192, dtype=torch.int32, device='cuda')
sink = torch.zeros(128*1024*1024, dtype=torch.float32, device='cuda')  # 512 MB > L2
print('timing build:', _lib.set_timing_buffer(buf))
B,H,NC,G = 1,48,32,16
d = O.make_inputs(B,H,NC,seed=1)
bf = lambda t: t.to(torch.bfloat16).cuda()
prm = [d[k].cuda().requires_grad_(True) for k in ('ln_w','ln_b','W1','b1','W2','b2')]
q,v,k = [bf(d[n]).requires_grad_(True) for n in ('XQ','XV','XK')]
e = bf(d['eta'])[:,:,:,-1,:].clone().requires_grad_(True)
go = bf(d['dOut'])
side = torch.cuda.Stream()
for label, mode, blocks, threads, smem in [('code28K 100x256', 9, 100, 256, 200*1024), ('code30K 100x256', 10, 100, 256, 200*1024), ('code32K 100x256', 11, 100, 256, 200*1024),
                                           ('code36K 100x256', 12, 100, 256, 200*1024), ('code24K 100x256', 6, 100, 256, 200*1024), ('none', -1, 0, 0, 0)]:
    for rep in range(2):
        out = mlp_tk.ttt_mlp_op(*prm, q, v, k, e, G)
        torch.cuda.synchronize()
        buf.zero_()
        if mode >= 0:
            rc = _lib.debug_lib().ttt_b200_debug_spin(blocks, threads, 6000000, mode, smem, _lib.ptr(sink), sink.numel(), side.cuda_stream)
            assert rc == 0, rc
        out.backward(go)
        torch.cuda.synchronize()
    t = buf.cpu().tolist()
    for gi in (1, 0):
        sts = [x & 0xffffffff for x in t[4096+gi*32:4096+gi*32+16]]
        print('%%-14s group %%d step us:' %% (label, gi), [round(((sts[kk-1]-sts[kk]) & 0xffffffff)/1e3,1) for kk in range(15,0,-1)], flush=True)
""" % (ROOT, ROOT, ROOT),
    "attn_bwd": """
import torch, sys, time
sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests')
from oracle import ttt_oracle as O
from ttt_video_dit_b200 import attention
for (B,T,H) in [(1,128,1),(2,300,3),(1,1000,2),(1,129,1),(1,4096,16)]:
    g = torch.Generator().manual_seed(1000 + T)
    q, k, v, go = (torch.randn(B, T, H, 64, generator=g).to(torch.bfloat16) for _ in range(4))
    q = q * 2.0
    qc, kc, vc = (t.cuda().requires_grad_(True) for t in (q, k, v))
    try:
        out = attention.sdpa_bthd(qc, kc, vc); out.backward(go.cuda()); torch.cuda.synchronize()
    except Exception as ex:
        print((B,T,H), 'EXC', repr(ex)[:300], flush=True); break
    qr, kr, vr = (t.float().requires_grad_(True) for t in (q, k, v))
    tr = lambda t: t.permute(0, 2, 1, 3)
    ref = O.sdpa_math(tr(qr), tr(kr), tr(vr)).permute(0, 2, 1, 3); ref.backward(go.float())
    print((B,T,H), 'out', '%%.2e' %% O.rel_err(out.float().cpu(), ref.detach()), {n: float('%%.2e' %% O.rel_err(a.grad.float().cpu(), b.grad)) for n,a,b in (('dq',qc,qr),('dk',kc,kr),('dv',vc,vr))}, flush=True)
    g1 = [t.grad.clone() for t in (qc,kc,vc)]
    same = True
    for rep in range(3):
        for t in (qc,kc,vc): t.grad = None
        out = attention.sdpa_bthd(qc, kc, vc); out.backward(go.cuda()); torch.cuda.synchronize()
        same = same and all(torch.equal(a, t.grad) for a, t in zip(g1, (qc,kc,vc)))
    print('   bitwise reproducible over 3 more runs:', same, flush=True)
B,T,H = 1,18048,48
q, k, v, go = (torch.randn(B, T, H, 64, device='cuda').to(torch.bfloat16) for _ in range(4))
qc, kc, vc = (t.requires_grad_(True) for t in (q, k, v))
import torch.nn.functional as F
for rep in range(2):
    torch.cuda.synchronize(); t0=time.time()
    out = attention.sdpa_bthd(qc, kc, vc); torch.cuda.synchronize(); t1=time.time()
    out.backward(go); torch.cuda.synchronize(); t2=time.time()
    print('ours  T=18048 H=48: fwd %%.2f ms bwd %%.2f ms (bwd %%.0f TF/s algorithmic)' %% ((t1-t0)*1e3, (t2-t1)*1e3, 2.5*4*T*T*64*H/(t2-t1)/1e12), flush=True)
qh, kh, vh = (t.detach().permute(0,2,1,3).requires_grad_(True) for t in (q,k,v))
for rep in range(2):
    torch.cuda.synchronize(); t0=time.time()
    o2 = F.scaled_dot_product_attention(qh, kh, vh); torch.cuda.synchronize(); t1=time.time()
    o2.backward(go.permute(0,2,1,3)); torch.cuda.synchronize(); t2=time.time()
    print('torch SDPA: fwd %%.2f ms bwd %%.2f ms' %% ((t1-t0)*1e3, (t2-t1)*1e3), flush=True)
""" % (ROOT, ROOT),
    "bwd_direct": """
import torch, sys
sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests')
from oracle import ttt_oracle as O
from ttt_video_dit_b200 import test_time_training as tt
from test_gpu_mlp_forward import run_forward
d = O.make_inputs(1,1,1,seed=41)
(q,k,v,le), out, ck, last = run_forward(d, 1)
go = d['dOut'].to(torch.bfloat16).cuda()
lw = d['ln_w'].float().cuda().reshape(1,1,1,64); lb = d['ln_b'].float().cuda().reshape(1,1,1,64)
try:
    r = tt.ttt_backward_simple(q,k,v,le,lw,lb,*ck,go,1)
    torch.cuda.synchronize()
    print('direct call ok', [float(x.float().abs().sum()) for x in r], flush=True)
except Exception as ex:
    print('direct EXC', repr(ex)[:600], flush=True)
""" % (ROOT, ROOT),
    "bwd": """
import torch, sys
sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests')
from oracle import ttt_oracle as O
from test_gpu_mlp_backward import errors
for (B,H,NC,G) in [(1,1,1,1),(1,1,2,1),(1,1,2,2),(1,2,3,2),(2,2,7,3),(1,3,20,16)]:
    d = O.make_inputs(B,H,NC,seed=40+NC)
    try:
        e = errors(d, G)
        print('bwd', (B,H,NC,G), {k: float('%%.3g' %% v) for k,v in e.items()}, flush=True)
    except Exception as ex:
        print('bwd', (B,H,NC,G), 'EXC', repr(ex)[:500], flush=True); break
""" % (ROOT, ROOT),
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(PROBES)
    for n in names:
        print(f"=== probe {n}", flush=True)
        try:
            r = subprocess.run([sys.executable, "-c", PROBES[n]], timeout=300, capture_output=True, text=True)
            print(r.stdout[-6000:])
            if r.returncode != 0:
                print("RC", r.returncode, r.stderr[-3000:])
        except subprocess.TimeoutExpired as e:
            print("TIMEOUT", (e.stdout or b"")[-3000:])
