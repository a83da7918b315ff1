"""Small end-to-end exercise of every kernel, meant to be run under compute-sanitizer:
    compute-sanitizer --tool memcheck  python tools/sanitize_run.py
    compute-sanitizer --tool racecheck python tools/sanitize_run.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fastkv_b200 import ops, _lib as L

g = torch.Generator(device="cuda").manual_seed(3)
def mk(b, l, h, d=128, dt=torch.bfloat16):
    return torch.randn(b, l, h, d, generator=g, device="cuda", dtype=dt).transpose(1, 2)

for (b, l, c, t, w, ks, pool) in [(1, 5000, 512, 1024, 8, 7, "avgpool"), (2, 1500, 128, 0, 8, 5, "maxpool"), (1, 3000, 256, 600, 16, 13, "avgpool")]:
    q, k, v = mk(b, l, 16), mk(b, l, 4), mk(b, l, 4)
    ops.compress(q, k, v, window_size=w, kernel_size=ks, pooling=pool, capacity=c, tsp_len=t)
    ops.compress(q, k, v, window_size=w, kernel_size=ks, pooling=pool, capacity=c, mode=L.MODE_SNAPKV)
    s = ops.score(q, k, window_size=1, kernel_size=5, mode=L.MODE_GEMFILTER)
    ops.topk(s.view(b, -1), 200, order=L.ORDER_INDEX_ASC)
    qf, kf, vf = q.float(), k.float(), v.float()
    ops.compress(qf, kf, vf, window_size=w, kernel_size=ks, pooling=pool, capacity=c, tsp_len=t)     # fp32 CUDA-core path
os.environ["FKV_SMALL_FUSE"] = "1"
q, k, v = mk(1, 2048, 32), mk(1, 2048, 8), mk(1, 2048, 8)
ops.compress(q, k, v, window_size=8, kernel_size=7, pooling="avgpool", capacity=2048)
os.environ["FKV_SMALL_FUSE"] = "0"
ops.compress(q, k, v, window_size=8, kernel_size=7, pooling="avgpool", capacity=2048)
# long rows: producer-built histogram, slice scan, per-row barrier, candidate list, crowded-bin refinement (narrow scores)
q, k, v = mk(1, 9000, 32), mk(1, 9000, 8), mk(1, 9000, 8)
ops.compress(q, k, v, window_size=8, kernel_size=7, pooling="avgpool", capacity=512, tsp_len=1024)
kq = (k * 0.01).contiguous().transpose(1, 2).contiguous().transpose(1, 2)            # flat scores: the threshold bin is crowded
ops.compress(q, kq, v, window_size=8, kernel_size=7, pooling="avgpool", capacity=2048)
ov = ops.SelectOverlap(partition=True)                                             # select + compact on a side stream, SM partition
for _ in range(3):
    ops.compress(q, k, v, window_size=8, kernel_size=7, pooling="avgpool", capacity=512, overlap=ov)
ov.join()
ops.topk(torch.rand(3, 70000, device="cuda"), 30000)
hid = torch.randn(1, 5000, 1024, device="cuda", dtype=torch.bfloat16)
ops.gather_rows(hid, torch.arange(0, 5000, 7, device="cuda")[None], torch.arange(5000, device="cuda")[None])
qd = torch.randn(2, 16, 128, device="cuda", dtype=torch.bfloat16)
kc = torch.randn(2, 4, 1024, 128, device="cuda", dtype=torch.bfloat16); vc = torch.randn_like(kc)
ops.cache_append(torch.randn(2, 4, 1, 128, device="cuda", dtype=torch.bfloat16), torch.randn(2, 4, 1, 128, device="cuda", dtype=torch.bfloat16), kc, vc, 900)
ops.decode_attn(qd, kc, vc, 901)                                                    # several splits + in-kernel merge
ops.decode_attn(qd, kc, vc, 40)                                                     # single CTA per head group
sl = torch.tensor([901], dtype=torch.int32, device="cuda")
ops.decode_attn(qd, kc, vc, 0, seq_len_dev=sl)
# the peer-memory score exchange kernels (one rank: the sanitizer serialises kernels, so ranks that wait for each other on
# one device cannot run under it; the flag protocol is covered by tests/test_gpu_exchange.py): vector path, scalar path
from fastkv_b200.peer import PeerRegion
for (hl, p_) in [(2, 4088), (3, 1021)]:
    (reg,) = PeerRegion.local_group(1, PeerRegion.bytes_for(2, hl, p_), torch.device("cuda", 0))
    for it in range(2):
        src = torch.rand(2, hl, p_, device="cuda")
        reg.push(src)
        out = reg.wait()
        torch.cuda.synchronize()
        assert torch.equal(out, src) and reg.status() == 0
torch.cuda.synchronize()
print("sanitize run done")
