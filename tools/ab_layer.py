"""A/B timing of one stride-1 tensor-core layer between two builds of libpccgeo on the same box:
    python tools/ab_layer.py tools/bin/libpccgeo_old.so pcc_geo_cnn_v2_b200/libpccgeo.so [B C S terms]"""
import ctypes as C
import sys

import numpy as np
import torch

paths = sys.argv[1:3]
B, Cc, S, terms = [int(a) for a in (sys.argv[3:] + ['32', '16', '64', '2'][len(sys.argv) - 3:])]
vp, i32, i64 = C.c_void_p, C.c_int, C.c_longlong
libs = []
for p in paths:
    l = C.CDLL(p)
    l.pccgeo_umma_pack_weights_host.restype = i64
    l.pccgeo_umma_pack_weights_host.argtypes = [vp, vp, i32, i32, i32, i32, i32]
    l.pccgeo_conv3d_umma.restype = i32
    l.pccgeo_conv3d_umma.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, i32, i32, vp]
    l.pccgeo_f32_to_blocked.argtypes = [vp, vp, i32, i32, i32, i32, i32, i32, vp]
    libs.append(l)
rng = np.random.default_rng(0)
x = torch.randn(B, Cc, S, S, S, device='cuda').relu_()
w = np.ascontiguousarray((rng.normal(size=(27, Cc, Cc)) / np.sqrt(27 * Cc)).astype(np.float32))
bias = torch.zeros(Cc, device='cuda')
xb = torch.empty(terms * x.numel(), device='cuda', dtype=torch.bfloat16)
libs[0].pccgeo_f32_to_blocked(x.data_ptr(), xb.data_ptr(), B, Cc, S, S, S, terms, None)
outs, wps = [], []
for l in libs:
    size = l.pccgeo_umma_pack_weights_host(w.ctypes.data, None, Cc, Cc, 1, 1, terms)
    img = np.zeros(size, np.uint8)
    l.pccgeo_umma_pack_weights_host(w.ctypes.data, img.ctypes.data, Cc, Cc, 1, 1, terms)
    wps.append(torch.from_numpy(img).cuda())
    outs.append(torch.empty_like(xb))


def run(i):
    rc = libs[i].pccgeo_conv3d_umma(xb.data_ptr(), wps[i].data_ptr(), bias.data_ptr(), None, outs[i].data_ptr(), B, Cc, S, S, S, Cc, 1, 1, 1,
                                    terms, None)
    assert rc == 0


for rnd in range(4):
    for i in range(len(libs)):
        for _ in range(3):
            run(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        print(f'round {rnd} {paths[i]}: {e0.elapsed_time(e1) / 20:.4f} ms', flush=True)
print('identical outputs:', torch.equal(outs[0], outs[1]))
