#!/usr/bin/env python
"""Offline check of the planner's cost model (capi.cu make_plan) against a per-kernel sweep (tools/size_sweep.py --all):
prints, per shape, the kernel the model picks, the measured best and the loss.  EFF mirrors g_kernels[].eff."""
import json, math, sys
path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/size_sweep.json"
d = json.load(open(path))
#            bm   bn  ctas/SM eff
CFG = {'128x128x16_s6': (128, 128, 1, 1.17), '128x128x32_s3': (128, 128, 1, 1.20), '128x64x32_s4': (128, 64, 1, 1.02),
       '96x64x32_s4': (96, 64, 1, 1.01), '64x64x64_s3': (64, 64, 1, 1.01), '96x64x32_s4_w8': (96, 64, 1, 1.17),
       '64x64x32_s3_x2': (64, 64, 2, 1.175), '128x64x32_s4_w8': (128, 64, 1, 1.185), '32x32x64_s3_x2': (32, 32, 2, 1.06),
       '64x32x32_s4_x2': (64, 32, 2, 1.125)}
for a in [x for x in sys.argv[2:] if "=" in x]:
    k, v = a.split("=")
    CFG[k] = CFG[k][:3] + (float(v),)

def model(M, N, n):
    bm, bn, c, eff = CFG[n]
    tiles = math.ceil(M / bm) * math.ceil(N / bn)
    res = 148 * c
    rounds = math.ceil(tiles / res)
    last = tiles - (rounds - 1) * res
    share = math.ceil(last / 148)
    waves = (rounds - 1) * c + (1.25 * share if share < c else share)
    return waves * bm * bn / eff

tot = 0
for k, v in d.items():
    if 'per_kernel_ms' not in v or not k.startswith('float64'):
        continue
    M, N, K = map(int, k.split('_')[1].split('x'))
    p = {n.replace('dmma_tma_f64_', ''): t for n, t in v['per_kernel_ms'].items()}
    pick = min(p, key=lambda n: model(M, N, n))
    best = min(p, key=p.get)
    loss = 100 * (p[pick] / p[best] - 1)
    tot += loss
    print(f"{k:28s} pick {pick:18s} {p[pick]:.4f}  best {best:18s} {p[best]:.4f}  loss {loss:5.1f}%  cublas {v['cublas_ms']:.4f}")
    if '-v' in sys.argv:
        print('    ' + ' '.join(f"{n}={t:.4f}" for n, t in p.items()))
print("total loss %.1f" % tot)
