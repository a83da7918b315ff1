import sys, os, torch
sys.path.insert(0, '/root/repo')
from fastkv_b200 import ops, _lib as L
for l in (32768, 131072):
    g = torch.Generator(device="cuda").manual_seed(0)
    nb = 3
    bufs = [(torch.randn(1, l, 32, 128, generator=g, device="cuda", dtype=torch.bfloat16).transpose(1, 2),
             torch.randn(1, l, 8, 128, generator=g, device="cuda", dtype=torch.bfloat16).transpose(1, 2)) for _ in range(nb)]
    for slot in (0, 1):
        def run():
            for q, k in bufs:
                ops.score(q, k, window_size=8, kernel_size=7, pooling="avgpool", hist_slot=slot)
        run(); torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            run()
        gr.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): gr.replay()
        e1.record(); torch.cuda.synchronize()
        print(f"L={l} hist_slot={slot}: {e0.elapsed_time(e1) * 1e3 / (10 * nb):.2f} us/call", flush=True)
    # leave the tables clean again
    ws = L.workspace(torch.device('cuda', 0), 1 << 24); L.check(L.load().fkv_workspace_init(ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream))
