import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import shgan_oracle as O
from shgan_b200 import kernels as K, packing as P
def t(a): return torch.from_numpy(np.ascontiguousarray(a)).cuda()
def run(x, w, impl, passes=3, block_n=0):
    xp = K.nchw_to_planes(t(x)); wh, wl = P.pack_conv_weight(t(w))
    n, _, h, wd = x.shape
    y = torch.empty((n, h, wd, w.shape[0]), device='cuda')
    K.conv_igemm([xp], wh, wl, P.taps_plain(3, 3), h, wd, epi=K.make_epilogue(out_f32=y), passes=passes, impl=impl, block_n=block_n)
    return K.nhwc_to_nchw_f32(y).cpu().numpy().astype(np.float64)
g = np.random.default_rng(0)
for (n, ci, co, h, w) in [(1,64,64,8,16),(1,128,64,8,16),(1,192,64,8,16),(1,512,64,8,16),(1,128,256,8,16),(2,512,512,16,16)]:
    x = g.standard_normal((n, ci, h, w)).astype(np.float32)
    wt = g.standard_normal((co, ci, 3, 3)).astype(np.float32)
    ref = O.conv2d(x.astype(np.float64), wt.astype(np.float64), padding=1)
    sc = np.abs(ref).max()
    y3, y1, ys = run(x, wt, 0, 3), run(x, wt, 0, 1), run(x, wt, 1)
    print(f'C{ci}->{co}: tc3 {np.abs(y3-ref).max()/sc:.2e}  tc1 {np.abs(y1-ref).max()/sc:.2e}  simt {np.abs(ys-ref).max()/sc:.2e}  tc3-simt {np.abs(y3-ys).max()/sc:.2e}')
    if ci > 64:
        for lohi, sl in (('first64', slice(0, 64)), ('rest', slice(64, None))):
            x2 = np.zeros_like(x); x2[:, sl] = x[:, sl]
            ref2 = O.conv2d(x2.astype(np.float64), wt.astype(np.float64), padding=1)
            print(f'   only {lohi}: tc3 {np.abs(run(x2, wt, 0, 3)-ref2).max()/sc:.2e} tc1 {np.abs(run(x2, wt, 0, 1)-ref2).max()/sc:.2e}')
        # precision probes: hi-only activations / hi-only weights
        xh = x.astype(np.float16).astype(np.float32); wh_ = wt.astype(np.float16).astype(np.float32)
        for nm, xx, ww in (('x=hi', xh, wt), ('w=hi', x, wh_), ('both hi', xh, wh_)):
            r = O.conv2d(xx.astype(np.float64), ww.astype(np.float64), padding=1)
            print(f'   tc3 vs oracle with {nm}: {np.abs(y3 - r).max()/sc:.2e}')

# end-to-end generator error vs golden for both conv implementations
import helpers as H
from golden.make_golden import GENERATOR_CASES
for name, res, chb, chm, batch, seed in GENERATOR_CASES:
    sd = O.synthetic_state_dict(res, seed=seed, ch_base=chb, ch_max=chm)
    G = H.build_generator(res, sd, chb, chm, device='cuda')
    x, z = O.synthetic_inputs(batch, res, seed=seed)
    gd = np.load(os.path.join(ROOT, 'tests', 'golden', name + '.npz'))
    outs = {}
    for impl in (1, 0):
        G.engine(impl=impl)
        outs[impl] = G(t(x), t(z), None, noise_mode='const').cpu().numpy().astype(np.float64)
        xg, _ = G.encoder(t(x))
        print(f'{name} impl={impl}: |img|max {np.abs(gd["img"]).max():.2f}  img max-abs err {np.abs(outs[impl]-gd["img"]).max():.3e}  x_global err {np.abs(xg.cpu().numpy()-gd["x_global"]).max():.3e}')
    print(f'{name}: tc vs simt {np.abs(outs[0]-outs[1]).max():.3e}')
