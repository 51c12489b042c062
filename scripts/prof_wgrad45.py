import os, sys, torch
sys.path.insert(0, '.')
from coinstac_dinunet_b200.ops import conv3d_wgrad as cw
dev = torch.device('cuda'); torch.manual_seed(0)
for cin, cout, shape in ((64, 128, (8, 15, 18, 15)), (128, 256, (8, 7, 9, 7))):
    x = torch.randn(*shape, cin, device=dev).bfloat16(); dy = torch.randn(*shape, cout, device=dev).bfloat16()
    buf = torch.zeros(27 * cin, cout, device=dev)
    for impl in ('tap', 'gather'):
        os.environ['COINN_WGRAD_IMPL'] = impl
        for _ in range(3): cw.conv3d_wgrad(dy, x, raw_out=buf)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): cw.conv3d_wgrad(dy, x, raw_out=buf)
        e1.record(); torch.cuda.synchronize()
        print(cin, cout, impl, cw.last_impl, round(e0.elapsed_time(e1) / 10 * 1e3, 1), 'us', flush=True)
