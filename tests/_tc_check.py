import sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from fastkv_b200 import ops, _lib as L
from parity_utils import make_inputs
torch.manual_seed(0)
for (b,hq,hkv,l,w,ks,pool) in [(1,32,8,2048,8,7,'avgpool'),(1,32,8,4099,8,7,'avgpool'),(2,16,4,2048,16,5,'maxpool'),(1,32,8,32768,8,7,'avgpool'),(1,8,8,1500,8,7,'avgpool')]:
    q,k,v = [x.cuda() for x in make_inputs(l, b,hq,hkv,l,128)]
    s_simple = ops.score(q,k,window_size=w,kernel_size=ks,pooling=pool,impl=L.IMPL_SIMPLE)
    torch.cuda.synchronize()
    s_tc = ops.score(q,k,window_size=w,kernel_size=ks,pooling=pool,impl=L.IMPL_TENSOR)
    torch.cuda.synchronize()
    ref = s_simple.double(); got = s_tc.double()
    scale = ref.abs().amax(-1, keepdim=True)
    err = ((ref-got).abs()/(ref.abs()+1e-3*scale)).max().item()
    print((b,hq,hkv,l,w,ks,pool), 'max rel err', err, 'nan', torch.isnan(s_tc).sum().item(), flush=True)
    if err > 1e-3:
        bad = ((ref-got).abs()/(ref.abs()+1e-3*scale) > 1e-3).nonzero()
        print('bad count', bad.shape[0], bad[:10].tolist(), bad[-5:].tolist())
        print(ref[tuple(bad[0])].item(), got[tuple(bad[0])].item())
