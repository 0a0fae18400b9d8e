#!/usr/bin/env python
"""Clock64 timeline of CTA 0 of the cta_group::1 FiLM-SIREN kernel (debug hook cng_internal_set_tc_trace).
    python tools/trace_tc.py [cta_group=1] [infer|train]      (train: the training-mode kernel with its dumps, 1 Mi points)"""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["CNG_TC_CG"] = sys.argv[1] if len(sys.argv) > 1 else "1"
import torch
from conditioned_nerf_gan_b200 import _lib, ops
from oracle import nerf_path as oracle
lib = _lib.load()
lib.cng_internal_set_tc_trace.argtypes = [ctypes.c_void_p]
lib.cng_internal_set_tc_trace.restype = None
MODE = sys.argv[2] if len(sys.argv) > 2 else "infer"
B, N, L = (8, 128 * 128 * 24, 8) if MODE == "infer" else (1, 1 << 20, 8)
st = oracle.init_generator_state("TALLSIREN_FG", seed=0)
dev = "cuda"
ws = [st[f"siren.network.{i}.layer.weight"].to(dev) for i in range(L)]
bs = [st[f"siren.network.{i}.layer.bias"].to(dev) for i in range(L)]
g = torch.Generator().manual_seed(1)
glob = torch.randn((B, 256), generator=g) * 0.05 + 0.19
freq, phase = (t.to(dev) for t in oracle.film_parameters(glob, st["siren.mapping_network.weight"], st["siren.mapping_network.bias"]))
feat = (torch.randn((B, N, 32), generator=g) * 0.3).to(dev)
fw, fb = st["siren.final_layer.weight"].to(dev), st["siren.final_layer.bias"].to(dev)
run = (lambda: ops.film_siren_fwd(feat, ws, bs, freq, phase, fw, fb, True, "bf16")) if MODE == "infer" else \
      (lambda: ops.film_siren_fwd_train(feat, ws, bs, freq, phase, fw, fb, True, "fp16"))
run(); torch.cuda.synchronize()
trace = torch.zeros((4, 9, 2, 8), dtype=torch.int64, device=dev)
lib.cng_internal_set_tc_trace(ctypes.c_void_p(trace.data_ptr()))
run(); torch.cuda.synchronize()
lib.cng_internal_set_tc_trace(None)
t = trace.cpu()
t0 = int(t[..., :4][t[..., :4] > 0].min())
print("iter layer slot | act_ready_seen  mma_issued | acc_full_seen  epi_done | epi_dur  act->acc_full  epi_done->next_act_seen | w_wait issue")
for it in range(1, 2):
    for l in range(9):
        for x in range(2):
            a, b, c, d = [int(v) - t0 if v > 0 else -1 for v in t[it, l, x, :4]]
            nxt = int(t[it, l + 1, x, 0]) - t0 if l < 8 else (int(t[it + 1, 0, x, 0]) - t0)
            print(f"{it:3d} {l:5d} {x:4d} | {a:10d} {b:10d} | {c:10d} {d:10d} | {d - c if d > 0 else -1:7d} {c - a if c > 0 else -1:9d} {nxt - d if d > 0 else -1:9d} | {int(t[it, l, x, 4]):6d} {int(t[it, l, x, 5]):6d}")

print("per-chunk stamps of iter 1, layer 3 (cycles from the slot-layer's act_ready): before wait, after wait, after 4 MMAs, after commit")
for x in range(2):
    base = int(t[1, 3, x, 0])
    for c in range(4):
        print(f"  slot {x} chunk {c}: ", [int(v) - base for v in t[3, c, x, :4]])
    print(f"  slot {x}: mma_issued {int(t[1, 3, x, 1]) - base}, acc_full seen by epilogue {int(t[1, 3, x, 2]) - base}")
