"""Fused generator-forward engine: the SH-GAN hot path as a fixed sequence of sm_100a kernel launches.

`GeneratorEngine` reads the parameters of a `model_zoo.comodgan.Generator` (reference layout), packs them once
into kernel operands (tap-major fp16 hi/lo weights, demodulation energy tables, SHU constants) and then runs
    mapping -> encoder (+SHU) -> style affines / demod coefficients -> synthesis (+ fused torgb / skip adds)
entirely through the C ABI.  Activations stay on the device in the split-plane NHWC layout between kernels;
NCHW fp32 exists only at the network input and the RGB output (reference call: comodgan.py:449-481).

Per-layer mapping of reference ops to launches:
  encoder conv0 / b4.conv   : 1 x shgan_conv_igemm (bias + lrelu_agc fused)
  encoder conv1 (down 2)    : shgan_fir_nhwc (blur -> 4 parity planes) + shgan_conv_igemm over the 4 planes
  synthesis conv0 (up 2)    : 4 x shgan_conv_igemm RAW parity passes (transposed conv at algorithmic cost)
                              + shgan_fir_nhwc (blur, demod, noise, bias, lrelu, +feats[res], next-layer style)
  synthesis conv1 / b4.conv : 1 x shgan_conv_igemm (demod, noise, bias, lrelu, fused torgb, next-layer style)
  torgb + img upsample      : shgan_torgb_combine
"""
import functools
import gc
import math

import torch

from . import kernels as K
from . import packing as P
from .kernels import Planes

SQRT2 = math.sqrt(2.0)


def _check_device(dev):
    if dev.type != 'cuda':
        raise RuntimeError('shgan_b200 runs on CUDA devices only (no CPU fallback): move the generator to cuda')


def _on_engine_device(fn):
    """Runs an engine entry point with the parameters' device as the current CUDA device, so that streams, graph capture,
    RNG state and every kernel launch go to the GPU that owns the model even when the caller never called
    torch.cuda.set_device (the reference eval does not: lib/experiments/shgan_default.py:164)."""
    @functools.wraps(fn)
    def wrapped(self, *args, **kwargs):
        self._ensure()
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, torch.Tensor) and a.device != self.dev:
                raise RuntimeError(f'{type(self).__name__}.{fn.__name__}: input on {a.device}, parameters on {self.dev}')
        if self.dev.type != 'cuda' or torch.cuda.current_device() == self.dev.index:    # (CPU: host-logic tests only)
            return fn(self, *args, **kwargs)
        with torch.cuda.device(self.dev):
            return fn(self, *args, **kwargs)
    return wrapped


class _Act:
    def __init__(self, spec):
        a = P.parse_activation(spec) if isinstance(spec, (str, type(None))) else spec
        self.on = a is not None
        self.alpha = a['alpha'] if a else 0.0
        self.gain = a['gain'] if a else 1.0
        self.clamp = (a['clamp'] if a and a['clamp'] is not None else -1.0)


class GeneratorEngine:
    """passes: 3 = fp32-class split-precision tensor-core convolutions (parity mode), 1 = single fp16 pass.
    impl: 0 = tcgen05 kernel (product), 1 = fp32 FMA cross-check kernel (tests only)."""

    def __init__(self, G, passes=3, impl=0, graphs=True):
        self.G = G
        self.passes = passes
        self.impl = impl
        self.graphs = graphs      # replay the ~120 launches of a forward as one CUDA graph per (shape, mode)
        self.fuse_up2 = True      # synthesis conv0: one shgan_conv_up2 launch instead of 4 RAW parity passes + the blur kernel
        self.overlap = True       # run the small off-critical-path kernels (mapping, SHU, torgb combine) on a side stream
        self._side = None
        self._sig = None
        self._buf = {}
        self._graphs = {}

    # ---- parameter packing ---------------------------------------------------------------------
    def _signature(self):
        sig = []
        for t in list(self.G.parameters()) + list(self.G.buffers()):
            sig.append((t.data_ptr(), t._version))
        return tuple(sig)

    def refresh(self):
        """(Re)pack every kernel operand from the module's current parameters."""
        G = self.G
        enc, syn = G.encoder, G.synthesis
        dev = next(G.parameters()).device
        _check_device(dev)
        if dev.type == 'cuda' and dev.index is None:
            dev = torch.device('cuda', torch.cuda.current_device())
        self.dev = dev
        self.res = syn.resolution
        self.act = _Act(syn.activation)
        f = syn.b4.conv.resample_filter if getattr(syn.b4.conv, 'resample_filter', None) is not None else None
        f = P.setup_filter([1, 3, 3, 1]) if f is None else f
        self.f = f.detach().to(dev, torch.float32).contiguous()               # as stored (upsample2d flips it itself)
        self.f_applied = self.f.flip([0, 1]).contiguous()                     # correlation taps of upfirdn2d(flip_filter=False)
        self.f_sep = P.separable_taps(self.f_applied)                         # (fy, fx) when the blur is rank 1 (it always is: [1,3,3,1])

        # mapping
        m = G.mapping
        self.map_layers = []
        for i in range(m.num_layers):
            fc = getattr(m, f'fc{i}')
            self.map_layers.append((fc.weight.detach(), fc.bias.detach() if fc.bias is not None else None,
                                    float(fc.weight_gain), float(fc.bias_gain), _Act(fc.activation_spec)))

        # encoder
        self.enc_res = list(enc.encode_res)
        self.enc = {}
        eact = _Act(enc.activation)
        self.eact = eact
        for idx, r in enumerate(self.enc_res[:-1]):
            b = getattr(enc, f'b{r}')
            d = {}
            if b.fromrgb is not None:
                w = b.fromrgb.weight.detach()
                d['fromrgb'] = (w.reshape(w.shape[0], w.shape[1]).contiguous(), b.fromrgb.bias.detach(),
                                float(b.fromrgb.weight_gain))
            for nm in ('conv0', 'conv1'):
                l = getattr(b, nm)
                wh, wl = P.pack_conv_weight(l.weight)
                d[nm] = dict(w_hi=wh, w_lo=wl, bias=l.bias.detach(), wgain=float(l.weight_gain),
                             ci=l.weight.shape[1], co=l.weight.shape[0])
            self.enc[r] = d
        l = enc.b4.conv
        wh, wl = P.pack_conv_weight(l.weight)
        self.enc[4] = dict(conv=dict(w_hi=wh, w_lo=wl, bias=l.bias.detach(), wgain=float(l.weight_gain),
                                     ci=l.weight.shape[1], co=l.weight.shape[0]),
                           fc=(enc.b4.fc.weight.detach(), enc.b4.fc.bias.detach(), float(enc.b4.fc.weight_gain),
                               float(enc.b4.fc.bias_gain), _Act(enc.b4.fc.activation_spec)))
        # SHU
        shu = getattr(enc, 'shu', None)
        self.shu = None
        if shu is not None:
            c2 = shu.in_channels * 2
            r_in, r_lo = shu.input_res, shu.lowest_res
            masks = P.gaussian_band_masks(r_in, r_lo, shu.tail_sigma_mult, shu.gaussian_at_input_res)
            self.shu = dict(
                ch=shu.in_channels, input_res=r_in, lowest_res=r_lo, reslist=sorted(masks),
                conv0_w=shu.conv0.weight.detach().reshape(c2, c2).contiguous(), conv0_b=shu.conv0.bias.detach(),
                df1_w=shu.df1.weight.detach().contiguous(),
                cw=P.make_cweight(shu.df1.freedom, (r_in, r_in // 2 + 1)).to(dev).contiguous(),
                gauss=torch.cat([masks[r].reshape(-1) for r in sorted(masks)]).to(dev).contiguous())
            # the channel mix's fp16 operands are packed here, once per parameter set, not on every forward
            self.shu['packed'] = K.shu_pack(self.shu['conv0_w'].float().contiguous(), self.shu['df1_w'].float().contiguous())

        # synthesis
        self.syn_res = list(syn.block_res)
        self.syn = {}
        self.style_layers = []     # (name, ws index, affine w, affine b, demod, wsq, pre_scale, ci, co)

        def syn_layer(l, name, widx):
            w_hat, wsq = P.demod_weight(l.weight)
            wh, wl = P.pack_conv_weight(w_hat)
            up2 = P.pack_up2_weight(w_hat) if (getattr(l, 'up', 1) == 2 and w_hat.shape[0] % 64 == 0 and w_hat.shape[2] == 3) else None
            self.style_layers.append(dict(name=name, widx=widx, aw=l.affine.weight.detach(), ab=l.affine.bias.detach(),
                                          again=float(l.affine.weight_gain), demod=True, wsq=wsq, pre_scale=1.0,
                                          ci=l.weight.shape[1], co=l.weight.shape[0]))
            return dict(w_hi=wh, w_lo=wl, bias=l.bias.detach(), noise_const=l.noise_const.detach().contiguous(),
                        noise_strength=l.noise_strength.detach(), ci=l.weight.shape[1], co=l.weight.shape[0],
                        res=l.resolution, use_noise=l.use_noise, up2=up2)

        def rgb_layer(l, name, widx):
            w = l.weight.detach()
            self.style_layers.append(dict(name=name, widx=widx, aw=l.affine.weight.detach(), ab=l.affine.bias.detach(),
                                          again=float(l.affine.weight_gain), demod=False, wsq=None,
                                          pre_scale=float(l.weight_gain), ci=w.shape[1], co=w.shape[0]))
            wpad = torch.zeros((3, w.shape[1]), dtype=torch.float32, device=dev)
            wpad[:w.shape[0]] = w.reshape(w.shape[0], w.shape[1])
            bias = torch.zeros(3, dtype=torch.float32, device=dev)
            bias[:w.shape[0]] = l.bias.detach()
            return dict(w=wpad.contiguous(), bias=bias)

        widx = 0
        b4 = syn.b4
        self.syn[4] = dict(fc=(b4.fc.weight.detach(), b4.fc.bias.detach(), float(b4.fc.weight_gain), float(b4.fc.bias_gain),
                               _Act(b4.fc.activation_spec)),
                           conv=syn_layer(b4.conv, 'b4.conv', widx), torgb=rgb_layer(b4.torgb, 'b4.torgb', widx + 1))
        widx += 1
        for r in self.syn_res[1:]:
            b = getattr(syn, f'b{r}')
            self.syn[r] = dict(conv0=syn_layer(b.conv0, f'b{r}.conv0', widx), conv1=syn_layer(b.conv1, f'b{r}.conv1', widx + 1),
                               torgb=rgb_layer(b.torgb, f'b{r}.torgb', widx + 2))
            widx += 2
        # every style affine reads the same input [w ; x_global] when ws is a broadcast w (always, on the eval path:
        # comodgan.py:449-481 repeats w to num_ws): one dense call over the concatenated weights serves all layers
        gains = {L['again'] for L in self.style_layers}
        self.aff_cat = None
        if len(gains) == 1 and len(self.style_layers) <= 40:
            off = 0
            for L in self.style_layers:
                L['offset'] = off
                off += L['ci']
            self.aff_cat = dict(w=torch.cat([L['aw'] for L in self.style_layers], dim=0).contiguous(),
                                b=torch.cat([L['ab'] for L in self.style_layers], dim=0).contiguous(),
                                gain=gains.pop(), total=off)
        self._sig = self._signature()

    def _ensure(self):
        if self._sig is None or self._sig != self._signature():
            self.refresh()
            self._buf = {}
            self._graphs = {}     # captured graphs reference the old packed operands

    # ---- side stream ---------------------------------------------------------------------------------
    class _Fork:
        """`with engine._fork():` runs the enclosed launches on the engine's side stream, ordered after everything already
        enqueued on the caller's stream; `engine._join()` makes the caller's stream wait for them.  Inside a CUDA-graph
        capture these become fork / join edges of the graph.  The kernels moved there are latency-bound launches that sit
        between two large convolutions of the main chain (mapping: 9 launches; the SHU: 4; torgb_combine: 8)."""

        def __init__(self, eng):
            self.eng = eng
            self.ctx = None

        def __enter__(self):
            if not self.eng.overlap or self.eng.dev.type != 'cuda':
                return self
            if self.eng._side is None:
                self.eng._side = torch.cuda.Stream(device=self.eng.dev)
            self.eng._side.wait_stream(torch.cuda.current_stream())
            self.ctx = torch.cuda.stream(self.eng._side)
            self.ctx.__enter__()
            return self

        def __exit__(self, *exc):
            if self.ctx is not None:
                self.ctx.__exit__(*exc)
            return False

    def _fork(self):
        return GeneratorEngine._Fork(self)

    def _join(self):
        if self.overlap and self._side is not None:
            torch.cuda.current_stream().wait_stream(self._side)

    # ---- buffers --------------------------------------------------------------------------------------
    def _planes(self, name, n, h, w, c):
        key = (name, n, h, w, c)
        b = self._buf.get(key)
        if b is None:
            b = Planes.empty(n, h, w, c, self.dev)
            self._buf[key] = b
        return b

    def _f32(self, name, *shape):
        key = (name,) + tuple(shape)
        b = self._buf.get(key)
        if b is None:
            b = torch.zeros(shape, dtype=torch.float32, device=self.dev)
            self._buf[key] = b
        return b

    # ---- pieces -----------------------------------------------------------------------------------------
    def _dense(self, spec, x0, out, x1=None):
        w, b, wg, bg, act = spec
        return K.dense(x0, w, b, out, wg, bg, act.on, act.alpha, act.gain, act.clamp, x1=x1)

    @_on_engine_device
    def mapping(self, z):
        """Mapping.forward (stylegan.py:394-430) for c_dim == 0, truncation_psi == 1 -> w [N, w_dim]."""
        self._ensure()
        n = z.shape[0]
        x = K.normalize_2nd_moment(z.contiguous().float(), self._f32('map_in', n, z.shape[1]))
        for i, spec in enumerate(self.map_layers):
            x = self._dense(spec, x, self._f32(f'map{i}', n, spec[0].shape[0]))
        return x

    SPLITK_MAX_PIXELS = 2048          # N*OH*OW up to here (4x4 / 8x8 at batch 16-32): too few tiles to fill 148 SMs
    SPLITK_BYTES = 16 << 20

    def _splitk_scratch(self, n, oh, ow):
        """fp32 scratch for the K-split of the 4x4 / 8x8 convolutions (shgan_conv_igemm, ACT mode with z set); None elsewhere."""
        if n * oh * ow > self.SPLITK_MAX_PIXELS:
            return None
        b = self._buf.get('splitk')
        if b is None:
            b = torch.empty(self.SPLITK_BYTES // 4, dtype=torch.float32, device=self.dev)
            self._buf['splitk'] = b
        return b

    def _conv(self, srcs, L, taps, oh, ow, epi=None, raw=None):
        sk = self._splitk_scratch(srcs[0].shape[0], oh, ow) if raw is None else None
        K.conv_igemm(srcs, L['w_hi'], L['w_lo'], taps, oh, ow, epi=epi, raw=raw, passes=self.passes, impl=self.impl, splitk=sk)

    def _enc_epi(self, L, out):
        a = self.eact
        return K.make_epilogue(wgain=L['wgain'], bias=L['bias'], act=a.on, act_alpha=a.alpha, act_gain=a.gain,
                               act_clamp=a.clamp, out=out)

    @_on_engine_device
    def encoder(self, x):
        """shgan.Encoder.forward (shgan.py:361-383) -> (x_global fp32 [N,oc_n], feats {res: Planes})."""
        self._ensure()
        masked = isinstance(x, (tuple, list))      # (real [N,3,R,R], mask [N,1,R,R]): input preparation fused into fromrgb
        if masked:
            real, mask = x[0].contiguous().float(), x[1].contiguous().float()
            n, _, R, _ = real.shape
            x = self._f32('x.prepared', n, real.shape[1] + 1, R, R)     # written by the fromrgb kernel, read by the composite
            self._prepared_x = x
        else:
            x = x.contiguous().float()
            n, _, R, _ = x.shape
        feats = {}
        a = None
        ident = None
        shu_outs = None
        for r in self.enc_res[:-1]:
            d = self.enc[r]
            c = d['conv0']['ci']
            if 'fromrgb' in d:
                w, b, wg = d['fromrgb']
                if masked:      # x = cat([mask - 0.5, real * mask]) (shgan_default.py:269-274) formed inside the kernel
                    a = K.fromrgb_masked(real, mask, x, w, b, wg, self.eact.alpha, self.eact.gain, self.eact.clamp,
                                         self._planes(f'e{r}.rgb', n, r, r, c))
                else:
                    a = K.fromrgb(x, w, b, wg, self.eact.alpha, self.eact.gain, self.eact.clamp, self._planes(f'e{r}.rgb', n, r, r, c))
            feat = self._planes(f'e{r}.feat', n, r, r, d['conv0']['co'])
            self._conv([a], d['conv0'], P.taps_plain(3, 3), r, r, epi=self._enc_epi(d['conv0'], feat))
            feats[r] = feat
            if self.shu is not None and r == self.shu['input_res']:
                with self._fork():                      # overlaps the rest of the encoder; its outputs are added after the join
                    shu_outs = self._shu_compute(feat, n)
            # conv1: blur (pad 2) into four parity planes, then the stride-2 conv as a 4-source stride-1 conv
            ph = (r + 1 + 1) // 2
            c1 = d['conv1']['ci']
            par = self._planes(f'e{r}.par', 4 * n, ph, ph, c1)
            ident = K.make_epilogue(out=par)
            K.fir_nhwc(feat, self.f_applied, 1.0, (2, 2, 2, 2), ident, parity_split=True, rank1=self.f_sep is not None)
            srcs = [Planes(par.hi[q * n:(q + 1) * n], par.lo[q * n:(q + 1) * n]) for q in range(4)]
            a = self._planes(f'e{r}.down', n, r // 2, r // 2, d['conv1']['co'])
            self._conv(srcs, d['conv1'], P.taps_down2(3), r // 2, r // 2, epi=self._enc_epi(d['conv1'], a))
        d = self.enc[4]
        feat4 = self._planes('e4.feat', n, 4, 4, d['conv']['co'])
        self._conv([a], d['conv'], P.taps_plain(3, 3), 4, 4, epi=self._enc_epi(d['conv'], feat4))
        feats[4] = feat4
        flat = K.planes_to_nchw(feat4, out=self._f32('e4.flat', n, d['conv']['co'], 4, 4))
        x_global = self._dense(d['fc'], flat.view(n, -1), self._f32('x_global', n, d['fc'][0].shape[0]))
        if self.shu is not None:
            if shu_outs is None:                        # input_res == 4: the SHU input is the last encoder feature
                shu_outs = self._shu_compute(feats[self.shu['input_res']], n)
            self._join()
            ch = self.shu['ch']
            rl = self.shu['reslist']
            # feats[r][:, -ch:] += shu[r] for every band (shgan.py:378-382), one launch
            K.planes_add_nchw_multi([feats[r] for r in rl], shu_outs, [feats[r].shape[3] - ch for r in rl])
        return x_global, feats

    def _shu_compute(self, fin, n):
        """SHU.forward (shgan.py:312-336) on the last `ch` channels of feats[input_res] -> list of fp32 [N,ch,r,r]."""
        s = self.shu
        ch, rin = s['ch'], s['input_res']
        xin = K.planes_to_nchw(fin, c_off=fin.shape[3] - ch, c=ch, out=self._f32('shu.in', n, ch, rin, rin))
        outs = [self._f32(f'shu.out{r}', n, ch, r, r) for r in s['reslist']]
        ws = self._buf.get(('shu.ws', n))
        if ws is None:
            ws = torch.empty(K.shu_workspace_bytes(n, ch, rin), dtype=torch.uint8, device=self.dev)
            self._buf[('shu.ws', n)] = ws
        K.shu_fwd(xin, s['conv0_w'], s['conv0_b'], s['df1_w'], s['cw'], s['gauss'], outs, s['lowest_res'], workspace=ws,
                  packed=s['packed'])
        return outs

    def styles(self, ws, x_global):
        """Affine transforms + style normalisation / demodulation coefficients for every synthesis layer
        (stylegan.py:280, 331 and :145-155).  ws [N,num_ws,w_dim] (any strides with unit inner stride)."""
        n = x_global.shape[0]
        out = {}
        if self.aff_cat is not None and ws.stride(1) == 0:
            A = self.aff_cat
            raw = self._f32('st.rawcat', n, A['total'])
            K.dense(ws[:, 0], A['w'], A['b'], raw, A['gain'], 1.0, False, x1=x_global)
            layers = []
            for L in self.style_layers:
                s_hat = self._f32('st.hat.' + L['name'], n, L['ci'])
                dcoef = self._f32('st.dc.' + L['name'], n, L['co']) if L['demod'] else None
                layers.append(dict(offset=L['offset'], ci=L['ci'], co=L['co'], demod=L['demod'], pre_scale=L['pre_scale'],
                                   wsq=L['wsq'], s_hat=s_hat, dcoef=dcoef))
                out[L['name']] = (s_hat, dcoef)
            K.style_prep_batched(raw, layers)
            return out
        for L in self.style_layers:
            raw = self._f32('st.raw.' + L['name'], n, L['ci'])
            K.dense(ws[:, L['widx']], L['aw'], L['ab'], raw, L['again'], 1.0, False, x1=x_global)
            s_hat = self._f32('st.hat.' + L['name'], n, L['ci'])
            dcoef = self._f32('st.dc.' + L['name'], n, L['co']) if L['demod'] else None
            K.style_prep(raw, L['wsq'], s_hat, dcoef, L['demod'], L['pre_scale'])
            out[L['name']] = (s_hat, dcoef)
        return out

    def _noise(self, L, n, noise_mode, gen=None):
        if not L['use_noise'] or noise_mode == 'none':
            return None, 0
        r = L['res']
        if noise_mode == 'const':
            return L['noise_const'], 0
        return torch.randn([n, 1, r, r], device=self.dev, generator=gen), r * r  # stylegan.py:282-283

    def draw_noise(self, n, noise_mode):
        """The per-layer noise inputs of one synthesis pass, drawn in the reference's layer order (stylegan.py:282-283:
        b4.conv, then conv0, conv1 of every block), so that a fixed torch seed gives the reference's noise.  Returned as
        {layer name: (tensor, per-sample stride)}; `_forward_eager` draws them on the side stream while the encoder runs."""
        out = {}
        for r in self.syn_res:
            d = self.syn[r]
            if r == 4:
                out['b4.conv'] = self._noise(d['conv'], n, noise_mode)
            else:
                out[f'b{r}.conv0'] = self._noise(d['conv0'], n, noise_mode)
                out[f'b{r}.conv1'] = self._noise(d['conv1'], n, noise_mode)
        return out

    @_on_engine_device
    def synthesis(self, x_global, feats, ws, noise_mode='random', comp_x=None, noise=None):
        """comodgan.Synthesis.forward (comodgan.py:396-433).  Returns img fp32 [N,3,R,R] (+ uint8 composite)."""
        self._ensure()
        n = x_global.shape[0]
        act = self.act
        if noise is None:
            noise = self.draw_noise(n, noise_mode)
        st = self.styles(ws, x_global)
        names = [L['name'] for L in self.style_layers]

        def next_conv_style(name):
            i = names.index(name) + 1
            while i < len(names) and names[i].endswith('torgb'):
                i += 1
            return st[names[i]][0] if i < len(names) else None

        d = self.syn[4]
        c4 = d['conv']['ci']
        x0 = self._dense(d['fc'], x_global, self._f32('s4.fc', n, d['fc'][0].shape[0]))
        x = K.nchw_to_planes(x0.view(n, c4, 4, 4), add=feats[4], scale=st['b4.conv'][0], out=self._planes('s4.in', n, 4, 4, c4))
        img = None
        last = self.syn_res[-1]
        comp = None
        for r in self.syn_res:
            d = self.syn[r]
            if r == 4:
                L, name = d['conv'], 'b4.conv'
            else:
                # conv0: stride-2 transposed conv as four parity passes at algorithmic cost, then blur + epilogue
                L0, name0 = d['conv0'], f'b{r}.conv0'
                h = r // 2
                nz, sn = noise[name0]
                y = self._planes(f's{r}.mid', n, r, r, L0['co'])
                epi = K.make_epilogue(dcoef=st[name0][1], noise=nz, noise_sn=sn, noise_strength=L0['noise_strength'],
                                      bias=L0['bias'], act=act.on, act_alpha=act.alpha, act_gain=act.gain, act_clamp=act.clamp,
                                      skip=feats[r], next_scale=st[f'b{r}.conv1'][0], out=y)
                if self.fuse_up2 and self.impl == 0 and L0['up2'] is not None and self.f_sep is not None:
                    # one launch: transposed conv (4 parities stacked along N) + blur + epilogue, z never leaves the SM
                    K.conv_up2(x, L0['up2'][0], L0['up2'][1], self.f_sep[1], self.f_sep[0], 4.0, epi, passes=self.passes)
                else:
                    z = self._f32(f's{r}.z', n, 2 * h + 1, 2 * h + 1, L0['co'])
                    for py in range(2):
                        for px in range(2):
                            self._conv([x], L0, P.taps_up2(py, px), P.up2_pass_size(h, py), P.up2_pass_size(h, px),
                                       raw=(z, 2, 2, py, px))
                    K.fir_nhwc(z, self.f_applied, 4.0, (1, 1, 1, 1), epi, rank1=self.f_sep is not None)
                x = y
                L, name = d['conv1'], f'b{r}.conv1'
            nz, sn = noise[name]
            nblk = K.conv_num_nblocks(L['co'])
            part = self._f32(f's{r}.rgb', n, r, r, nblk, 4)
            nxt = next_conv_style(name)
            out = self._planes(f's{r}.out', n, r, r, L['co']) if r != last else None
            epi = K.make_epilogue(dcoef=st[name][1], noise=nz, noise_sn=sn, noise_strength=L['noise_strength'], bias=L['bias'],
                                  act=act.on, act_alpha=act.alpha, act_gain=act.gain, act_clamp=act.clamp,
                                  rgb_w=d['torgb']['w'], rgb_style=st[f'b{r}.torgb'][0], rgb_out=part,
                                  next_scale=nxt if out is not None else None, out=out)
            self._conv([x], L, P.taps_plain(3, 3), r, r, epi=epi)
            img_r = self._f32(f's{r}.img', n, 3, r, r)
            if r == last and comp_x is not None:
                comp = torch.empty((n, 3, r, r), dtype=torch.uint8, device=self.dev)
            with self._fork():                          # the image chain only meets the feature chain again at the very end
                K.torgb_combine(img, part, d['torgb']['bias'], self.f, img_r, comp_x=comp_x if r == last else None,
                                comp_out=comp if r == last else None)
            img = img_r
            x = out
        self._join()
        return (img, comp) if comp_x is not None else img

    @_on_engine_device
    def forward(self, x, z, noise_mode='random', composite=False):
        """comodgan.Generator.forward (comodgan.py:449-481).  composite=True additionally returns the eval loop's
        uint8 composite (shgan_default.py:257-262) fused into the last kernel.  With `graphs` the launch sequence is
        captured once per (shape, noise_mode, composite, passes, impl) and replayed; torch's graph-safe Philox keeps
        noise_mode='random' drawing fresh noise on every replay.  The returned tensors are engine-owned buffers."""
        self._ensure()
        masked = isinstance(x, (tuple, list))      # (real, mask): see encoder()
        xin = list(x) if masked else [x]
        if xin[0].shape[0] == 0:
            # empty batch: empty results, like the reference's PyTorch ops (the kernels take no null / zero-size tensors)
            r = xin[0].shape[-1]
            img = torch.empty((0, 3, r, r), dtype=torch.float32, device=xin[0].device)
            return (img, torch.empty((0, 3, r, r), dtype=torch.uint8, device=xin[0].device)) if composite else img
        if not (self.graphs and xin[0].is_cuda):
            return self._forward_eager(x, z, noise_mode, composite)
        key = (tuple(tuple(t.shape) for t in xin), tuple(z.shape), noise_mode, composite, self.passes, self.impl, self.fuse_up2)
        entry = self._graphs.get(key)
        if entry is None:
            xs = [t.detach().contiguous().float().clone() for t in xin]
            zs = z.detach().contiguous().float().clone()
            cur = torch.cuda.current_stream(self.dev)
            side = torch.cuda.Stream(device=self.dev)
            side.wait_stream(cur)
            rng = torch.cuda.get_rng_state(self.dev)
            xarg = tuple(xs) if masked else xs[0]
            with torch.cuda.stream(side):      # warm-up: allocates the persistent buffers, loads the kernels
                self._forward_eager(xarg, zs, noise_mode, composite)
            cur.wait_stream(side)
            torch.cuda.synchronize(self.dev)
            torch.cuda.set_rng_state(rng, self.dev)   # the warm-up must not consume the caller's random stream
            graph = torch.cuda.CUDAGraph()
            # no cyclic garbage collection while the stream is capturing: collecting an unreachable generator/engine of an
            # earlier call frees CUDA resources (graphs, side-stream blocks), and that invalidates the capture in progress
            gc_was_enabled = gc.isenabled()
            gc.disable()
            try:
                with torch.cuda.graph(graph):
                    out = self._forward_eager(xarg, zs, noise_mode, composite)
            finally:
                if gc_was_enabled:
                    gc.enable()
            entry = (graph, xs, zs, out)
            self._graphs[key] = entry
        graph, xs, zs, out = entry
        for dst, src in zip(xs, xin):
            dst.copy_(src, non_blocking=True)
        zs.copy_(z, non_blocking=True)
        graph.replay()
        return out

    @_on_engine_device
    def _forward_eager(self, x, z, noise_mode='random', composite=False):
        self._ensure()
        with self._fork():                              # the mapping network only meets the encoder at the style affines,
            w = self.mapping(z)                         # the noise inputs only meet the synthesis layers
            noise = self.draw_noise(z.shape[0], noise_mode)
        if self.dev.type == 'cuda' and not torch.cuda.is_current_stream_capturing():
            for t, _ in noise.values():                 # drawn on the side stream, consumed on this one
                if t is not None and noise_mode == 'random':
                    t.record_stream(torch.cuda.current_stream())
        num_ws = self.G.num_ws
        ws = w.unsqueeze(1).expand(w.shape[0], num_ws, w.shape[1])
        if not isinstance(x, (tuple, list)):
            x = x.contiguous().float()
        x_global, feats = self.encoder(x)
        if isinstance(x, (tuple, list)):
            x = self._prepared_x          # the 4-channel network input the fromrgb kernel wrote
        self._join()
        return self.synthesis(x_global, feats, ws, noise_mode=noise_mode, comp_x=x if composite else None, noise=noise)


class DiscriminatorEngine:
    """Discriminator.forward (stylegan.py:828-838) as a launch sequence over the same kernels: per block
    skip = 1x1 conv(downsample2d(x)) * sqrt(1/2) ; conv0 3x3 ; conv1 = blur + 3x3 stride 2, gain sqrt(1/2), + skip
    (discrim_block.forward :658-684), then minibatch-std -> conv 3x3 -> fc -> out (discrim_epilogue.forward :743-755)."""

    def __init__(self, D, passes=3, impl=0):
        self.D, self.passes, self.impl = D, passes, impl
        self._sig = None
        self._buf = {}

    def _signature(self):
        return tuple((t.data_ptr(), t._version) for t in list(self.D.parameters()) + list(self.D.buffers()))

    def refresh(self):
        D = self.D
        dev = next(D.parameters()).device
        _check_device(dev)
        if dev.type == 'cuda' and dev.index is None:
            dev = torch.device('cuda', torch.cuda.current_device())
        self.dev = dev
        self.act = _Act(D.activation)
        self.res_list = list(D.encode_res)
        f = getattr(D, f'b{self.res_list[0]}').resample_filter
        self.f_applied = f.detach().to(dev, torch.float32).flip([0, 1]).contiguous()
        self.f_rank1 = P.separable_taps(self.f_applied) is not None           # decided once on the host: no fallback launch, row-walking blur
        self.blocks = {}

        def conv(l, pad_ci=None):
            w = l.weight.detach().float()
            if pad_ci is not None and pad_ci > w.shape[1]:
                wp = torch.zeros((w.shape[0], pad_ci, w.shape[2], w.shape[3]), dtype=torch.float32, device=dev)
                wp[:, :w.shape[1]] = w
                w = wp
            wh, wl = P.pack_conv_weight(w)
            return dict(w_hi=wh, w_lo=wl, bias=None if l.bias is None else l.bias.detach(), wgain=float(l.weight_gain),
                        ci=w.shape[1], co=w.shape[0])
        for r in self.res_list[:-1]:
            b = getattr(D, f'b{r}')
            d = dict(conv0=conv(b.conv0), conv1=conv(b.conv1), skip=conv(b.skip))
            if b.fromrgb is not None:
                w = b.fromrgb.weight.detach()
                d['fromrgb'] = (w.reshape(w.shape[0], w.shape[1]).contiguous(), b.fromrgb.bias.detach(), float(b.fromrgb.weight_gain))
            self.blocks[r] = d
        b4 = D.b4
        c4 = b4.conv.weight.shape[0]
        self.mbstd = None if b4.mbstd is None else (b4.mbstd.group_size, b4.mbstd.num_channels)
        if self.mbstd is not None and self.mbstd[1] != 1:
            raise NotImplementedError('minibatch-std with more than one statistic channel')
        self.c4 = c4
        self.c4_pad = (b4.conv.weight.shape[1] + 63) // 64 * 64
        self.b4 = dict(conv=conv(b4.conv, self.c4_pad),
                       fc=(b4.fc.weight.detach(), b4.fc.bias.detach(), float(b4.fc.weight_gain), float(b4.fc.bias_gain), _Act(b4.fc.activation_spec)),
                       out=(b4.out.weight.detach(), b4.out.bias.detach(), float(b4.out.weight_gain), float(b4.out.bias_gain), _Act(b4.out.activation_spec)))
        self._sig = self._signature()

    def _planes(self, name, n, h, w, c):
        key = (name, n, h, w, c)
        if key not in self._buf:
            self._buf[key] = Planes.empty(n, h, w, c, self.dev)
        return self._buf[key]

    def _epi(self, L, out, gain=1.0, act=True, skip=None):
        a = self.act
        if act and a.on:
            return K.make_epilogue(wgain=L['wgain'], bias=L['bias'], act=True, act_alpha=a.alpha, act_gain=a.gain * gain,
                                   act_clamp=a.clamp * gain if a.clamp > 0 else -1.0, skip=skip, out=out)
        return K.make_epilogue(wgain=L['wgain'], bias=L['bias'], act=False, act_gain=gain, skip=skip, out=out)

    def _ensure(self):
        if self._sig is None or self._sig != self._signature():
            self.refresh()
            self._buf = {}

    @_on_engine_device
    def forward(self, img):
        img = img.contiguous().float()
        n = img.shape[0]
        if n == 0:
            return torch.empty((0, 1), dtype=torch.float32, device=img.device)      # empty batch: empty logits
        g = math.sqrt(0.5)
        conv = lambda srcs, L, taps, oh, ow, epi: K.conv_igemm(srcs, L['w_hi'], L['w_lo'], taps, oh, ow, epi=epi,
                                                               passes=self.passes, impl=self.impl)
        x = None
        for r in self.res_list[:-1]:
            d = self.blocks[r]
            c, cn = d['conv0']['ci'], d['conv1']['co']
            if 'fromrgb' in d:
                w, b, wg = d['fromrgb']
                x = K.fromrgb(img, w, b, wg, self.act.alpha, self.act.gain, self.act.clamp, self._planes(f'd{r}.rgb', n, r, r, c))
            # skip: downsample2d (blur pad 1, keep even samples) -> 1x1 conv, no bias / activation, gain sqrt(1/2)
            ds = self._planes(f'd{r}.ds', n, r // 2, r // 2, c)
            K.fir_nhwc(x, self.f_applied, 1.0, (1, 1, 1, 1), K.make_epilogue(out=ds), parity_split=2, rank1=self.f_rank1)
            y = self._planes(f'd{r}.skip', n, r // 2, r // 2, cn)
            conv([ds], d['skip'], P.taps_plain(1, 1), r // 2, r // 2, self._epi(d['skip'], y, gain=g, act=False))
            # conv0, then conv1 = blur (pad 2) into parity planes + stride-2 conv over them, gain sqrt(1/2), + skip
            t0 = self._planes(f'd{r}.c0', n, r, r, c)
            conv([x], d['conv0'], P.taps_plain(3, 3), r, r, self._epi(d['conv0'], t0))
            ph = (r + 2) // 2
            par = self._planes(f'd{r}.par', 4 * n, ph, ph, c)
            K.fir_nhwc(t0, self.f_applied, 1.0, (2, 2, 2, 2), K.make_epilogue(out=par), parity_split=1, rank1=self.f_rank1)
            srcs = [Planes(par.hi[q * n:(q + 1) * n], par.lo[q * n:(q + 1) * n]) for q in range(4)]
            x = self._planes(f'd{r}.out', n, r // 2, r // 2, cn)
            conv(srcs, d['conv1'], P.taps_down2(3), r // 2, r // 2, self._epi(d['conv1'], x, gain=g, skip=y))
        if self.mbstd is not None:
            xin = K.mbstd_append(x, self._planes('d4.mb', n, 4, 4, self.c4_pad), self.mbstd[0])
        else:
            xin = x
        t4 = self._planes('d4.conv', n, 4, 4, self.c4)
        conv([xin], self.b4['conv'], P.taps_plain(3, 3), 4, 4, self._epi(self.b4['conv'], t4))
        flat = K.planes_to_nchw(t4).view(n, -1)
        spec = self.b4['fc']
        h = torch.empty((n, spec[0].shape[0]), dtype=torch.float32, device=self.dev)
        K.dense(flat, spec[0], spec[1], h, spec[2], spec[3], spec[4].on, spec[4].alpha, spec[4].gain, spec[4].clamp)
        spec = self.b4['out']
        out = torch.empty((n, spec[0].shape[0]), dtype=torch.float32, device=self.dev)
        K.dense(h, spec[0], spec[1], out, spec[2], spec[3], spec[4].on, spec[4].alpha, spec[4].gain, spec[4].clamp)
        return out
