"""Development aid: fused first contraction (modes 2,3) on a mode-0 slab of `rows` rows of the order-4 s=300 tensor,
for different K splits (PPX_KSPLIT) -- the shapes the ranks of a multi-GPU run see."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
ppx = importlib.import_module("pairwise-perturbation_b200")
import torch
ctx = ppx.Ctx(0, workspace_bytes=int(os.environ.get("WS_GB", "2")) << 30)
R = 50
W = [ctx.empty(300 * R) for _ in range(4)]
for i, w in enumerate(W): ctx.fill_uniform(w, 2, i)
rows_list = [int(a) for a in sys.argv[1].split(",")] if len(sys.argv) > 1 else [37, 38]
splits = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
for rows in rows_list:
    lens = [rows, 300, 300, 300]
    P = rows * 300**3
    V = ctx.empty(P); ctx.fill_uniform(V, 1, 0)
    out = ctx.empty(rows * 300 * R)
    for S in splits:
        if S: os.environ["PPX_KSPLIT"] = str(S)
        ts = []
        for rep in range(4):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(ctx.stream):
                e0.record(); ctx.ttm_multi(V, lens, 2, [W[2], W[3]], R, out); e1.record()
            e1.synchronize(); ts.append(e0.elapsed_time(e1))
        os.environ.pop("PPX_KSPLIT", None)
        print(f"rows {rows} L={rows*300} split {S or 'auto'}: {min(ts):.3f} ms  {2*P*R/min(ts)/1e9:.1f} TF/s", flush=True)
    del V, out
