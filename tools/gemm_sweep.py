"""GPU: correctness + timing of the tcgen05 GEMM for every tiling / cluster shape on the headline shapes.
Each configuration runs in its own subprocess (a protocol bug then costs one line, not the whole sweep)."""
import os
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, ctypes, torch
sys.path.insert(0, %r)
import imagecaptioning.pytorch_b200 as b200
L = b200._lib; lib = L.load()
mode = L.MODES[sys.argv[1]]
SH = os.environ.get('SHAPES')
shapes = [tuple(int(v) for v in t.split('x')) for t in SH.split(',')] if SH else [(1280, 4000, 2000), (1280, 4000, 3000), (1280, 9488, 1000), (9216, 1000, 2048), (1280, 512, 1000), (256, 4000, 2000)]
for (M, N, K) in shapes:
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g); w = (torch.rand(N, K, generator=g) * 2 - 1) / K ** 0.5; b = torch.randn(N, generator=g)
    ref = x.double() @ w.double().t() + b.double()
    xd, wd, bd = x.cuda(), w.cuda(), b.cuda()
    y = torch.full((M, N), float('nan'), device='cuda')
    ms = ctypes.c_float(0)
    rc = lib.capb200_bench_linear(L.ptr(xd), L.ptr(wd), L.ptr(bd), L.ptr(y), M, N, K, mode, 20, ctypes.byref(ms), L.current_stream())
    err = float((y.cpu().double() - ref).abs().max()) if rc == 0 else float('nan')
    if rc != 0:
        lib.capb200_last_error.restype = ctypes.c_char_p
        print('  error:', lib.capb200_last_error().decode(), flush=True)
    print('  M=%%5d N=%%5d K=%%5d rc=%%d  %%.1f us  %%.0f TFLOP/s alg  max|err|=%%.2e' %% (M, N, K, rc, ms.value * 1e3, 2.0 * M * N * K / (ms.value * 1e-3) / 1e12 if ms.value > 0 else 0, err), flush=True)
''' % REPO

TILINGS = sys.argv[1:] or ['128x1x1', '144x1x1', '128x2x1', '144x2x1', '128x1x2', '144x1x2', '128x2x2', '144x2x2', 'pair128', 'pair144', 'pair192', 'pair256', 'auto']
for mode in ['tc_f16x3', 'tc_f16x1']:
    for tiling in TILINGS:
        env = dict(os.environ)
        if tiling != 'auto':
            env['CAPB200_GEMM_TILING'] = tiling
        else:
            env.pop('CAPB200_GEMM_TILING', None)
        print('== %s tiling %s' % (mode, tiling), flush=True)
        try:
            r = subprocess.run([sys.executable, '-c', CHILD, mode], env=env, capture_output=True, text=True, timeout=120)
            print(r.stdout, end='')
            if r.returncode != 0:
                print('  FAILED rc=%d: %s' % (r.returncode, r.stderr[-400:]))
        except subprocess.TimeoutExpired:
            print('  TIMEOUT')
