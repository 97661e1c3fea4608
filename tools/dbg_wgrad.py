import sys, torch, torch.nn.functional as F
from yolov5_obb_b200.train_ops import Wgrad, nhwc_to_nchw
DEV = "cuda:0"
def run(B, Cin, Cout, H, W, k, s):
    p = k // 2
    Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    g = torch.Generator().manual_seed(Cin + Cout)
    x = torch.randn(B, H, W, Cin, generator=g).bfloat16().to(DEV)
    dz = torch.randn(B, Ho, Wo, Cout, generator=g).bfloat16().to(DEV)
    xt = torch.zeros((B, Cin, H, W), dtype=torch.bfloat16, device=DEV)
    dzt = torch.zeros((B, Cout, Ho, Wo), dtype=torch.bfloat16, device=DEV)
    nhwc_to_nchw(x.data_ptr(), Cin, xt, B, Cin, H, W, phase_split=(s == 2))
    nhwc_to_nchw(dz.data_ptr(), Cout, dzt, B, Cout, Ho, Wo)
    torch.cuda.synchronize()
    dw = torch.zeros((k * k, Cout, Cin), dtype=torch.float32, device=DEV)
    wg = Wgrad(dzt, xt, dw, B, Cout, Ho, Wo, Cin, H, W, k, s, p)
    wg.run()
    torch.cuda.synchronize()
    w = torch.zeros((Cout, Cin, k, k), device=DEV, requires_grad=True)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), w, None, s, p)
    y.backward(dz.float().permute(0, 3, 1, 2))
    ref = w.grad.permute(2, 3, 0, 1).reshape(k * k, Cout, Cin)
    print((B, Cin, Cout, H, W, k, s), "err", (dw - ref).abs().max().item(), "scale", ref.abs().max().item(), flush=True)
case = [int(v) for v in sys.argv[1:8]]
run(*case)
