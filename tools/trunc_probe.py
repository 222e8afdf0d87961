"""GPU probe: how much of the tensor-core 3-pass GEMM error is a systematic (round-toward-zero accumulate) scale bias?"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import imagecaptioning.pytorch_b200 as b200
L = b200._lib; lib = L.load()
for K in (256, 512, 1000, 2000, 3000):
    M, N = 512, 640
    g = torch.Generator().manual_seed(K)
    for dist in ('randn', 'positive'):
        x = torch.randn(M, K, generator=g)
        w = (torch.rand(N, K, generator=g) * 2 - 1) / K ** 0.5
        if dist == 'positive':
            x = x.abs(); w = w.abs()
        ref = x.double() @ w.double().t()
        xd, wd = x.cuda(), w.cuda()              # keep the device copies alive for the asynchronous call
        for mode in ('tc_f16x3', 'simt_fp32'):
            y = torch.empty(M, N, device='cuda')
            L.check(lib.capb200_linear(L.ptr(xd), K, L.ptr(wd), K, None, L.ptr(y), N, M, N, K, 0, L.MODES[mode], L.current_stream()), 'linear')
            torch.cuda.synchronize()
            yd = y.cpu().double()
            sel = ref.abs() > 0.25 * ref.abs().max()
            rel = ((yd - ref) / ref)[sel]
            scale = float((yd * ref).sum() / (ref * ref).sum())
            resid = (yd - scale * ref)
            n_acc = 3 * ((K + 63) // 64) * 4
            print('K=%4d %-8s %-9s mean rel err %+.3e (per accumulate %+.3e = %.2f x 2^-24)  max|err| %.2e  max|err after rescale| %.2e' % (
                K, dist, mode, float(rel.mean()), float(rel.mean()) / n_acc, float(rel.mean()) / n_acc / 2 ** -24, float((yd - ref).abs().max()), float(resid.abs().max())))
