import sys, torch
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from fastkv_b200 import ops
from parity_utils import make_inputs
ok = True
ov = ops.SelectOverlap()
for (l, c, t) in [(9000, 512, 1024), (2048, 2048, 0), (5000, 256, 0), (9000, 512, 0), (33000, 2048, 2048)]:
    q, k, v = [x.cuda() for x in make_inputs(l, 1, 32, 8, l, 128)]
    a = ops.compress(q, k, v, window_size=8, kernel_size=7, pooling="avgpool", capacity=c, tsp_len=t, return_aux=True)
    b = ops.compress(q, k, v, window_size=8, kernel_size=7, pooling="avgpool", capacity=c, tsp_len=t, return_aux=True, overlap=ov)
    ov.join(); torch.cuda.synchronize()
    same = torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[3]["idx"], b[3]["idx"]) and (a[2] is None or torch.equal(a[2], b[2]))
    print(l, c, t, "same" if same else "DIFFERENT"); ok &= same
print("overlap ok" if ok else "overlap FAILED")
