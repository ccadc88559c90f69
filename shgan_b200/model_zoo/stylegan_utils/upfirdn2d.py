"""Module-path-compatible home of the upfirdn2d operator family (reference: lib/model_zoo/stylegan_utils/upfirdn2d.py).
The implementation is the sm_100a kernel behind the C ABI (shgan_upfirdn2d_fwd); there is no plugin JIT step and
no `_upfirdn2d_ref` fallback."""
from ...ops import setup_filter, upfirdn2d, filter2d, upsample2d, downsample2d  # noqa: F401
