"""Shared test helpers: reference-config builders for the shgan_b200 model zoo."""
import numpy as np
import torch

ACT = 'lrelu_agc(alpha=0.2, gain=sqrt_2, clamp=256)'


def generator_cfg(res, ch_base=32768, ch_max=512, num_ws=None):
    """The `shgan_g256` / `shgan_g512` model configs of the reference (configs/model/{shgan,comodgan,stylegan}.yaml)
    as plain dicts; ch_base/ch_max can be shrunk for small test models."""
    if num_ws is None:
        num_ws = {256: 14, 512: 16, 1024: 18}.get(res, 2 * int(np.log2(res)) - 2)
    m = dict(type='comodgan_mapping', args=dict(z_dim=512, c_dim=0, w_dim=512, num_ws=num_ws, num_layers=8, embed_features=None,
                                                layer_features=None, activation=ACT, lr_multiplier=0.01, w_avg_beta=0.995))
    e = dict(type='shgan_encoder', args=dict(resolution=res, ic_n=4, oc_n=1024, ch_base=ch_base, ch_max=ch_max,
                                             use_fp16_before_res=None, resample_filter=[1, 3, 3, 1], activation=ACT,
                                             mbstd_group_size=0, mbstd_c_n=0, c_dim=None, cmap_dim=None, use_dropout=True,
                                             has_extra_final_layer=False, shu_channels=32, shu_df_freedom=[2, 3],
                                             shu_df_type='piecewise_linear', shu_input_res=64, shu_lowest_res=4,
                                             shu_tail_sigma_mult=3, shu_gaussian_at_input_res=False))
    s = dict(type='comodgan_synthesis', args=dict(w_dim=512, w0_dim=1024, resolution=res, rgb_n=3, ch_base=ch_base, ch_max=ch_max,
                                                  use_fp16_after_res=None, resample_filter=[1, 3, 3, 1], activation=ACT))
    return dict(type='comodgan_generator', args=dict(mapping=m, encoder=e, synthesis=s))


def build_generator(res, sd, ch_base=32768, ch_max=512, device='cpu'):
    from shgan_b200.model_zoo import get_model
    G = get_model()(generator_cfg(res, ch_base, ch_max))
    if not hasattr(G.synthesis, 'num_ws'):
        pass
    G.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    return G.eval().requires_grad_(False).to(device)


def build_discriminator(res, sd, ch_base=32768, ch_max=512, device='cpu'):
    """`comodgan_discriminator` with the args of configs/model/comodgan.yaml:51-58."""
    from shgan_b200.model_zoo import get_model
    D = get_model()(dict(type='comodgan_discriminator', args=dict(ic_n=4, ch_base=ch_base, ch_max=ch_max, resolution=res,
                                                                  use_fp16_before_res=None)))
    D.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
    return D.eval().requires_grad_(False).to(device)
